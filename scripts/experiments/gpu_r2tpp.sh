mkdir -p gpurun_out/r2tpp
for tpp in 2 3 4; do for k in 20 100; do
DFW_WIDE_TPP=$tpp timeout 300 python bench.py --steps $k --warmup 5 --no-cpu-baseline > gpurun_out/r2tpp/bench_${tpp}_$k.json 2> gpurun_out/r2tpp/bench_${tpp}_$k.err; python -c "
import json
d=json.load(open('gpurun_out/r2tpp/bench_${tpp}_$k.json')); print('BENCH tpp=$tpp K=$k', round(d['value']/1e6,1), d['ms_per_step'], 'e2e', round(d['e2e']['value']/1e6,1), {k:v['ms'] for k,v in d['roofline']['stages'].items() if 'alone' in k})"; done; done
