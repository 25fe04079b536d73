#!/bin/bash
# The headline configuration (configs[1]) on the model-specialised library: metrics pass, default bench line with the CPU baseline, reference
# arm, ncu launch list of the bench command, and one --set full capture of the free-flight kernel (a quarter of the SMs busy).
#   gpurun --timeout 1800 -- 'bash tools/record_headline.sh r3'
TAG=${1:-r3}
bash tools/record_configs.sh ${TAG}_spec "1" > gpurun_out/${TAG}_headline_metrics.log 2>&1; tail -n 2 gpurun_out/${TAG}_headline_metrics.log | cut -c1-220
timeout 900 python bench.py > gpurun_out/${TAG}_bench_config1.json 2> gpurun_out/${TAG}_bench_config1.err; cut -c1-300 gpurun_out/${TAG}_bench_config1.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err; cut -c1-200 gpurun_out/${TAG}_bench_reference.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_bench_steps2_warmup1.csv python bench.py --steps 2 --warmup 1 --skip-cpu > gpurun_out/${TAG}_launches.log 2>&1; tail -n 1 gpurun_out/${TAG}_launches.log | cut -c1-120
MLB_SPECIALIZE=1 MLB_EVEN_SPLIT=0 MLB_BLOCK_SIZE=1024 timeout 600 ncu --set full --clock-control none --import-source on -k regex:flightLanes -c 1 -o gpurun_out/${TAG}_spec_full -f python tools/sweep.py --config 1 --lanes 37888 --blocks 1024 > gpurun_out/${TAG}_full.log 2>&1; tail -n 1 gpurun_out/${TAG}_full.log | cut -c1-160
