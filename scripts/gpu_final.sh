# Final round-2 validation: GPU tests, smoke, the driver's bench protocol for both arms, and the other workloads' lines.
mkdir -p gpurun_out/final
timeout 900 python -m pytest tests/ -m gpu -x -q > gpurun_out/final/pytest_gpu.log 2>&1; tail -2 gpurun_out/final/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" > gpurun_out/final/smoke.log 2>&1; tail -1 gpurun_out/final/smoke.log
timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/final/bench_reference.json 2> gpurun_out/final/bench_reference.err; cut -c1-300 gpurun_out/final/bench_reference.json
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/final/bench_k20.json 2> gpurun_out/final/bench_k20.err
timeout 600 python bench.py --gpus 1 --steps 100 --warmup 5 > gpurun_out/final/bench_k100.json 2> gpurun_out/final/bench_k100.err
for w in criteo_pruned criteo_qr twitter; do timeout 600 python bench.py --workload $w --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/final/bench_$w.json 2> gpurun_out/final/bench_$w.err; done
timeout 600 python bench.py --precision bf16 --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/final/bench_bf16.json 2> gpurun_out/final/bench_bf16.err
timeout 300 python scripts/e2e_stress.py 150 64 > gpurun_out/final/stress.log 2>&1; tail -1 gpurun_out/final/stress.log
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/final/bench_*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        st = d.get('roofline', {}).get('stages', {}) if d.get('roofline') else {}
        print(f.split('/')[-1], round(d['value'] / 1e6, 2), d.get('ms_per_step'), 'e2e', round(d['e2e']['value'] / 1e6, 2), 'frac', (d.get('roofline') or {}).get('frac'),
              'parity', (d.get('parity') or {}).get('max_rel'), 'b65536', {k: v.get('samples_per_s') for k, v in st.items() if '65536' in k}, (d.get('reference_cuda') or {}).get('value'))
    except Exception as e:
        print(f, 'ERR', e)
PY
