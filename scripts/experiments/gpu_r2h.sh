N=${1:-2}
mkdir -p gpurun_out/r2h
for cap in 111 74 48; do for w in criteo; do
DFW_PULL_ROWS_CTAS=$cap timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29562 bench.py --gpus $N --steps 20 --warmup 5 --workload $w > gpurun_out/r2h/bench_${w}_${N}gpu_cap$cap.json 2> gpurun_out/r2h/bench_${w}_${N}gpu_cap$cap.err; python -c "
import json
d=json.loads(open('gpurun_out/r2h/bench_${w}_${N}gpu_cap$cap.json').read().strip().splitlines()[-1]); print('BENCH cap $cap $w $N gpus', round(d['value']/1e6,1), d['ms_per_step'], 'e2e', round(d['e2e']['value']/1e6,1), {k:v['ms'] for k,v in d['roofline']['stages'].items()})"
done; done
