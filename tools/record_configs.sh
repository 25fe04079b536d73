#!/bin/bash
# One JSON line per BASELINE configuration (and the flight-log variants of the headline one) into gpurun_out/; copy them to profiles/ afterwards.
#   gpurun --timeout 2400 -- 'bash tools/record_configs.sh'
for c in 0 2 3 4; do
    python bench.py --config $c --steps 2 --warmup 1 --skip-cpu > gpurun_out/r2_bench_config$c.json 2> gpurun_out/r2_bench_config$c.err
done
python bench.py --config 1 --steps 2 --warmup 1 --skip-cpu --log-every 8 --log-rows 256 > gpurun_out/r2_bench_config1_flightlog.json 2> gpurun_out/r2_bench_config1_flightlog.err
python bench.py --config 1 --steps 2 --warmup 1 --skip-cpu --log-every 1 --log-rows 64 > gpurun_out/r2_bench_config1_flightlog_every1.json 2>> gpurun_out/r2_bench_config1_flightlog.err
python - <<PY
import json
for f in ("config0", "config2", "config3", "config4", "config1_flightlog", "config1_flightlog_every1"):
    try:
        d = json.loads(open("gpurun_out/r2_bench_%s.json" % f).read().strip().splitlines()[-1])
        print(f, "value %.4g e2e %.4g ms %.1f frac %.3f flight_ms %.1f lanes %d ok %d sims/s %.0f" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["frac"],
              d["roofline"]["kernel_ms"], d["config"]["lanes_total"], d["config"]["lanes_ok"], d["sims_per_sec"]), d.get("log_roofline", {}).get("achieved"), d.get("log_roofline", {}).get("rows_kept"))
    except Exception as e:
        print(f, "failed", e)
PY
tail -c 400 gpurun_out/r2_bench_config1_flightlog.err
