#!/bin/bash
# Round-end GPU visit: parity tests, bench (both arms), launch list, full ncu captures of the
# dominant kernels of config 3 and of the sparse-visibility (config 4) Schur kernels, smoke.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.csv 2>&1
nproc > gpurun_out/nproc.txt
timeout 600 python -m pytest tests -q -m gpu -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
for k in ${KERNELS:-k_eval5}; do
  timeout 250 ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -f -o gpurun_out/prof_$k python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_$k.log 2>&1
done
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_pair_frames|k_pair_blocks|k_schur_pairs|k_reduce_pairs|k_reduce_s|k_backsub|k_solve|k_eval5|k_view_blocks|k_post_eval' -c 120 --csv --log-file gpurun_out/launches_cfg4.csv python tools/stress_cfg4.py --frames 40000 --timed-iterations 3 --out gpurun_out/stress_cfg4_ncu.json > gpurun_out/ncu_cfg4.log 2>&1
for k in k_schur_pairs2 k_pair_blocks; do
  timeout 200 ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -f -o gpurun_out/prof_cfg4_$k python tools/stress_cfg4.py --frames 40000 --timed-iterations 3 --out gpurun_out/stress_cfg4_ncu.json > gpurun_out/ncu_cfg4_$k.log 2>&1
done
timeout 200 python tools/stress_cfg4.py --frames 100000 --out gpurun_out/stress_cfg4_final.json > gpurun_out/stress_cfg4_final.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
grep -v "^\[init\|Ceres Solver" gpurun_out/pytest_gpu.log | tail -4; tail -2 gpurun_out/smoke.log; head -c 600 gpurun_out/bench.json; echo; tail -2 gpurun_out/bench.err; head -c 300 gpurun_out/bench_ref.json; echo; tail -1 gpurun_out/stress_cfg4_final.log | head -c 600
