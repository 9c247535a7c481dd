set -x
mkdir -p gpurun_out/r2d
for d in 0 4 8 12 2; do DFW_WIDE_DBG=$d timeout 60 python scripts/wide_timeline.py 4096 bf16x3 > gpurun_out/r2d/tl4k_dbg$d.txt 2>&1; echo "== dbg $d"; grep "back-to-back\|MMA:\|Error\|error" gpurun_out/r2d/tl4k_dbg$d.txt | head -8; done
