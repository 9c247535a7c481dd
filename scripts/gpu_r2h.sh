N=${1:-2}
mkdir -p gpurun_out/r2h
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29561 tests/mp_sharded_check.py > gpurun_out/r2h/mp_sharded_check_${N}gpu.log 2>&1; echo "mp_sharded_check rc=$?"
for w in criteo twitter; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29562 bench.py --gpus $N --steps 100 --warmup 5 --workload $w > gpurun_out/r2h/bench_${w}_${N}gpu_pull.json 2> gpurun_out/r2h/bench_${w}_${N}gpu_pull.err; tail -2 gpurun_out/r2h/bench_${w}_${N}gpu_pull.err; python -c "
import json
d=json.loads(open('gpurun_out/r2h/bench_${w}_${N}gpu_pull.json').read().strip().splitlines()[-1]); print('BENCH pull $w $N gpus', round(d['value']/1e6,1), d['ms_per_step'], 'e2e', round(d['e2e']['value']/1e6,1), d['roofline']['stages'])"
done
for w in criteo_qr; do
timeout 600 python bench.py --steps 100 --warmup 5 --workload $w --no-cpu-baseline > gpurun_out/r2h/bench_${w}_1gpu.json 2> gpurun_out/r2h/bench_${w}_1gpu.err; tail -2 gpurun_out/r2h/bench_${w}_1gpu.err; python -c "
import json
d=json.loads(open('gpurun_out/r2h/bench_${w}_1gpu.json').read().strip().splitlines()[-1]); print('BENCH $w 1 gpu', round(d['value']/1e6,1), d['ms_per_step'], 'e2e', round(d['e2e']['value']/1e6,1), d['roofline']['stages'])"
done
