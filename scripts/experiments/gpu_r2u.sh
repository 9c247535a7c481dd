N=${1:-8}
mkdir -p gpurun_out/r2u
for ex in p2p; do for w in criteo twitter; do
DFW_BENCH_EXCHANGE=$ex timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29562 bench.py --gpus $N --steps 20 --warmup 5 --workload $w --no-cpu-baseline > gpurun_out/r2u/bench_${w}_${N}gpu_$ex.json 2> gpurun_out/r2u/bench_${w}_${N}gpu_$ex.err; python -c "
import json
d=json.loads(open('gpurun_out/r2u/bench_${w}_${N}gpu_$ex.json').read().strip().splitlines()[-1]); print('BENCH $ex $w $N gpus', round(d['value']/1e6,1), d['ms_per_step'], 'e2e', round(d['e2e']['value']/1e6,1), {k:v['ms'] for k,v in d['roofline']['stages'].items()})"
done; done
