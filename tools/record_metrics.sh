#!/bin/bash
# ncu metrics pass (one free-flight launch at full occupancy) of every BASELINE configuration: the files bench.py reads `roofline.traffic`
# and the executed FP64 instruction counts from (profiles/r2_config<C>_full_occupancy_metrics_lanes<N>.csv).
#   gpurun --timeout 1800 -- 'bash tools/record_metrics.sh'
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,l1tex__t_sector_pipe_lsu_mem_local_op_ld_hit_rate.pct,l1tex__t_sector_pipe_lsu_mem_local_op_st_hit_rate.pct,lts__t_sector_hit_rate.pct,sm__icc_request_hit_rate.pct,sm__inst_executed.avg.per_cycle_active,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__inst_executed.sum,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio
for spec in 0:151552 2:151552 3:100000 4:151552; do
    c=${spec%%:*}; n=${spec##*:}
    ncu --metrics $M --clock-control none -k regex:flightLanes -c 1 --csv --log-file gpurun_out/r2_config${c}_full_occupancy_metrics_lanes${n}.csv python tools/sweep.py --config $c --lanes $n > gpurun_out/r2_metrics_config$c.log 2>&1
    tail -n 1 gpurun_out/r2_metrics_config$c.log | cut -c1-200
done
