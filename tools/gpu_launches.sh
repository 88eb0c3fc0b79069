#!/bin/bash
# ncu launch list of the bench (per-launch durations; cold-cache, serialised)
mkdir -p gpurun_out
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-masked --stress-frames 0 > gpurun_out/ncu_bench.log 2>&1
python tools/launch_summary.py gpurun_out/launches.csv
