#!/bin/bash
# One JSON line per BASELINE configuration into gpurun_out/ (copy them to profiles/ afterwards), then the ncu metrics pass (one free-flight
# launch at full occupancy) of each: the files bench.py reads `roofline.traffic` and the executed FP64 instruction counts from.
#   gpurun --timeout 2400 -- 'bash tools/record_configs.sh r3_spec "0 2 3 4"'            (model-specialised libraries, the default)
#   gpurun --timeout 2400 -- 'bash tools/record_configs.sh r3_generic "1" --generic'
#   BLOCK=512 gpurun ... 'BLOCK=512 bash tools/record_configs.sh r3_spec "1"'            (metrics pass of the 512-lane CTA shape: file *_block512_*)
TAG=${1:-r3_spec}; CONFIGS=${2:-"0 2 3 4"}; FLAG=$3
SHAPE=""; SWEEPARGS=""; [ -n "$BLOCK" ] && { SHAPE="_block${BLOCK}"; SWEEPARGS="--blocks $BLOCK"; }
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,l1tex__t_sector_pipe_lsu_mem_local_op_ld_hit_rate.pct,l1tex__t_sector_pipe_lsu_mem_local_op_st_hit_rate.pct,lts__t_sector_hit_rate.pct,sm__icc_request_hit_rate.pct,sm__inst_executed.avg.per_cycle_active,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__inst_executed.sum,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio
for c in $CONFIGS; do
    n=151552; [ $c = 3 ] && n=100000
    if [ "$FLAG" = "--generic" ]; then export MLB_SPECIALIZE=0; else export MLB_SPECIALIZE=1; fi
    timeout 600 ncu --metrics $M --clock-control none -k regex:flightLanes -c 1 --csv --log-file gpurun_out/${TAG}_config${c}${SHAPE}_full_occupancy_metrics_lanes${n}.csv python tools/sweep.py --config $c --lanes $n $SWEEPARGS > gpurun_out/${TAG}_metrics_config$c.log 2>&1
    tail -n 1 gpurun_out/${TAG}_metrics_config$c.log | cut -c1-200
    cp gpurun_out/${TAG}_config${c}${SHAPE}_full_occupancy_metrics_lanes${n}.csv profiles/ 2>/dev/null      # bench.py reads it from profiles/ in this same call
    timeout 900 python bench.py --config $c --steps 2 --warmup 3 --skip-cpu $FLAG > gpurun_out/${TAG}_bench_config$c.json 2> gpurun_out/${TAG}_bench_config$c.err
    python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${TAG}_bench_config$c.json").read().strip().splitlines()[-1])
    print("config $c", "value %.4g e2e %.4g ms %.1f frac %.3f flight_ms %.1f lanes %d ok %d sims/s %.0f traffic/lane %.3g" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["frac"],
          d["roofline"]["kernel_ms"], d["config"]["lanes_total"], d["config"]["lanes_ok"], d["sims_per_sec"], (d["roofline"]["traffic"] or 0) / d["config"]["lanes_total"]))
except Exception as e:
    print("config $c failed", e)
PY
done
