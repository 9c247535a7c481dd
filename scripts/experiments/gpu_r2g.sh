N=${1:-2}
mkdir -p gpurun_out/r2g
nvidia-smi -L | head -8
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29561 tests/mp_sharded_check.py > gpurun_out/r2g/mp_sharded_check_${N}gpu.log 2>&1; echo "mp_sharded_check rc=$?"; grep -c "bit_identical=True" gpurun_out/r2g/mp_sharded_check_${N}gpu.log; grep "bit_identical=False" gpurun_out/r2g/mp_sharded_check_${N}gpu.log | head -5; tail -3 gpurun_out/r2g/mp_sharded_check_${N}gpu.log
for w in criteo twitter; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29562 bench.py --gpus $N --steps 100 --warmup 5 --workload $w > gpurun_out/r2g/bench_${w}_${N}gpu.json 2> gpurun_out/r2g/bench_${w}_${N}gpu.err; tail -2 gpurun_out/r2g/bench_${w}_${N}gpu.err; python -c "
import json
d=json.loads(open('gpurun_out/r2g/bench_${w}_${N}gpu.json').read().strip().splitlines()[-1]); print('BENCH $w $N gpus', round(d['value']/1e6,1), d['ms_per_step'], 'e2e', round(d['e2e']['value']/1e6,1), 'parity', d['parity'])"
done
