mkdir -p gpurun_out/final
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/final/bench_reference.json 2> gpurun_out/final/bench_reference.err
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/final/bench_k20.json 2> gpurun_out/final/bench_k20.err; tail -2 gpurun_out/final/bench_k20.err | grep -i error
timeout 600 python bench.py --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/final/bench_k100.json 2> gpurun_out/final/bench_k100.err
for w in criteo_pruned criteo_qr twitter; do timeout 600 python bench.py --steps 100 --warmup 5 --workload $w --no-cpu-baseline > gpurun_out/final/bench_${w}.json 2> gpurun_out/final/bench_${w}.err; done
timeout 300 python bench.py --steps 100 --warmup 5 --precision bf16 --no-cpu-baseline > gpurun_out/final/bench_bf16.json 2> gpurun_out/final/bench_bf16.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/final/bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, 'ERR', e); continue
    st=d.get('roofline',{}).get('stages',{})
    print(f.split('/')[-1], round(d['value']/1e6,2), d.get('ms_per_step'), 'e2e', round(d['e2e']['value']/1e6,1), 'i32', d['e2e'].get('int32_indices',{}).get('value'), 'frac', d.get('roofline',{}).get('frac'), 'big', st.get('fused_forward_batch_65536',{}).get('samples_per_s'), 'alone', st.get('fused_forward_one_launch_alone',{}).get('ms'), 'par', d.get('parity',{}).get('max_rel'), 'cpu', (d.get('cpu_baseline') or {}).get('value'), 'refcuda', (d.get('reference_cuda') or {}).get('value'))
PY
