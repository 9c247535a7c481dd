mkdir -p gpurun_out/r2l
timeout 900 python -m pytest tests -x -q -m gpu -k "fp32 or tiny or gathered or ragged or empty or non_contig or index_out or sigmoid" > gpurun_out/r2l/tests.log 2>&1; tail -3 gpurun_out/r2l/tests.log
timeout 600 python scripts/embed_roofline.py > gpurun_out/r2l/embed_roofline.json 2> gpurun_out/r2l/embed_roofline.err; tail -2 gpurun_out/r2l/embed_roofline.err; python -c "
import json
d=json.load(open('gpurun_out/r2l/embed_roofline.json'))
for r in d['runs']: print(r)"
