#!/bin/bash
# Profile visit: parity tests, bench, full ncu captures of the named kernels.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py --no-cpu-baseline --no-masked --stress-frames 0 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
for k in $KERNELS; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -f -o gpurun_out/prof_$k python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-masked --stress-frames 0 > gpurun_out/ncu_$k.log 2>&1
done
tail -8 gpurun_out/pytest_gpu.log; python - <<'PY'
import json
try:
    d = json.load(open('gpurun_out/bench.json'))
    print({k: d[k] for k in ('value','ms_per_step','lm_iterations_per_sec','stage_ms','gpu_launches')})
    print('k_eval ms', d['roofline']['ms_per_launch'], 'fp64 frac', d['roofline_fp64']['frac'], 'e2e', d['e2e']['lm_iterations_per_sec'], d['e2e']['seconds_per_call'])
except Exception as e:
    print('bench parse failed', e); print(open('gpurun_out/bench.err').read()[-2000:])
PY
