mkdir -p gpurun_out/r2j
timeout 600 python -m pytest tests -x -q -m gpu -k "qr or QR or golden_logits_bf16x3 or config4 or int32" > gpurun_out/r2j/tests.log 2>&1; tail -2 gpurun_out/r2j/tests.log
for w in criteo_qr; do timeout 120 python scripts/wide_timeline.py 4096 bf16x3 $w > gpurun_out/r2j/tl_$w.txt 2>&1; echo "== $w"; grep -v "epi:" gpurun_out/r2j/tl_$w.txt | head -9 | grep -v "^B="; done
