#!/bin/bash
# full ncu captures (with source) of the kernels matching regex $K from the bench ($N launches, default 1)
mkdir -p gpurun_out
timeout 500 ncu --set full --clock-control none --import-source on -k regex:$K -s ${SKIP:-2} -c ${N:-1} -f -o gpurun_out/prof_${OUT:-$K} python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_${OUT:-$K}.log 2>&1
tail -3 gpurun_out/ncu_${OUT:-$K}.log | cut -c1-300
