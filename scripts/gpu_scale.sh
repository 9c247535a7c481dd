mkdir -p gpurun_out/scale
nvidia-smi -L | wc -l
for N in 8 4; do for w in criteo twitter; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29562 bench.py --gpus $N --steps 20 --warmup 5 --workload $w > gpurun_out/scale/bench_${w}_${N}gpu.json 2> gpurun_out/scale/bench_${w}_${N}gpu.err; tail -2 gpurun_out/scale/bench_${w}_${N}gpu.err | grep -i error; python -c "
import json
d=json.loads(open('gpurun_out/scale/bench_${w}_${N}gpu.json').read().strip().splitlines()[-1]); print('BENCH $w $N gpus', round(d['value']/1e6,1), d['ms_per_step'], 'e2e', round(d['e2e']['value']/1e6,1), 'parity', d['parity']['max_rel'], d['parity'].get('ranks'), {k:v['ms'] for k,v in d['roofline']['stages'].items()})"
done; done
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29561 tests/mp_sharded_check.py > gpurun_out/scale/mp_sharded_check_8gpu.log 2>&1; echo "mp_sharded_check 8 rc=$?"; grep -c "bit_identical=True" gpurun_out/scale/mp_sharded_check_8gpu.log; grep -c "bit_identical=False" gpurun_out/scale/mp_sharded_check_8gpu.log
