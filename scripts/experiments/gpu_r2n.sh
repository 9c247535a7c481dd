mkdir -p gpurun_out/r2n
for tpp in 1 2 3 4; do for s in 8 16 24; do for k in 20 100; do
DFW_WIDE_TPP=$tpp timeout 200 python bench.py --steps $k --warmup 5 --no-cpu-baseline --streams $s > gpurun_out/r2n/b_${tpp}_${s}_${k}.json 2> gpurun_out/r2n/b_${tpp}_${s}_${k}.err; python -c "
import json
d=json.load(open('gpurun_out/r2n/b_${tpp}_${s}_${k}.json')); print('tpp $tpp streams $s K $k:', round(d['value']/1e6,1), d['ms_per_step'], 'e2e', round(d['e2e']['value']/1e6,1), 'alone', d['roofline']['stages']['fused_forward_one_launch_alone']['ms'])"
done; done; done
