#!/bin/bash
# Iteration visit: smoke (short timeout), parity tests, bench of the default path, k_solve cycle split.
mkdir -p gpurun_out
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; rc=$?; echo "smoke rc=$rc" >> gpurun_out/smoke.log
tail -2 gpurun_out/smoke.log
if [ $rc -ne 0 ]; then exit 1; fi
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
python - <<PY
import json
try:
    d = json.load(open('gpurun_out/bench.json'))
    print({k: d[k] for k in ('value','ms_per_step','lm_iterations_per_sec','stage_ms')})
    print('  k_eval ms', d['roofline']['ms_per_launch'], 'fp64 frac', d['roofline_fp64']['frac'], 'e2e', d['e2e']['lm_iterations_per_sec'])
except Exception as e:
    print('bench parse failed', e); print(open('gpurun_out/bench.err').read()[-2000:])
PY
timeout 120 bash tools/gpu_dbg.sh 2>&1 | grep "k_solve" | tail -1
