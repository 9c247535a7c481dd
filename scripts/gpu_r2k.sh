mkdir -p gpurun_out/r2k
for w in criteo criteo_qr_small; do
timeout 300 ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum,smsp__inst_executed_op_local_ld.sum,smsp__inst_executed_op_local_st.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct,smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct,smsp__warp_issue_stalled_no_instruction_per_warp_active.pct -k regex:fused_wide -s 2 -c 1 --csv --log-file gpurun_out/r2k/inst_$w.csv python scripts/run_one.py $w > gpurun_out/r2k/run_$w.log 2>&1
echo "== $w"; grep "fused_wide" gpurun_out/r2k/inst_$w.csv | awk -F'","' '{print $(NF-2), $(NF)}'
done
