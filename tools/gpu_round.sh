#!/bin/bash
# One GPU-box visit: parity tests, bench (both arms), launch list, full ncu captures.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.csv 2>&1
nproc > gpurun_out/nproc.txt
timeout 600 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-masked --stress-frames 0 > gpurun_out/ncu_bench.log 2>&1
for k in ${KERNELS:-k_eval5 k_view_blocks k_solve k_schur_frames k_schur_update k_backsub k_post_eval k_reduce_s}; do
  timeout 250 ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -f -o gpurun_out/prof_$k python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-masked --stress-frames 0 > gpurun_out/ncu_$k.log 2>&1
done
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
tail -3 gpurun_out/pytest_gpu.log; tail -2 gpurun_out/smoke.log; cat gpurun_out/bench.json | head -c 700; echo; tail -2 gpurun_out/bench.err; head -c 400 gpurun_out/bench_ref.json
