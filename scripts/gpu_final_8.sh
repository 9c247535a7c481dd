N=8
mkdir -p gpurun_out/finalN
for w in criteo twitter; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29562 bench.py --gpus $N --steps 20 --warmup 5 --workload $w --no-cpu-baseline > gpurun_out/finalN/bench_${w}_${N}gpu.json 2> gpurun_out/finalN/bench_${w}_${N}gpu.err; python -c "
import json
d=json.loads(open('gpurun_out/finalN/bench_${w}_${N}gpu.json').read().strip().splitlines()[-1]); print('BENCH $w $N gpus', round(d['value']/1e6,1), d['ms_per_step'], 'e2e', round(d['e2e']['value']/1e6,1), d['parity'].get('ok'), d['parity'].get('max_rel'), {k:v['ms'] for k,v in d['roofline']['stages'].items()})"
done
