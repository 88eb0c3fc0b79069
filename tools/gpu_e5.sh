#!/bin/bash
# k_eval5 bring-up: smoke under a short timeout first (a pipeline deadlock must not hang the box),
# then parity tests and the bench for both evaluation variants.
mkdir -p gpurun_out
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; rc=$?; echo "smoke rc=$rc" >> gpurun_out/smoke.log
tail -3 gpurun_out/smoke.log
if [ $rc -ne 0 ]; then exit 1; fi
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
for v in 5 4; do
  TSCM_EVAL_VARIANT=$v timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench_v$v.json 2> gpurun_out/bench_v$v.err; echo "bench rc=$?" >> gpurun_out/bench_v$v.err
  python - <<PY
import json
try:
    d = json.load(open('gpurun_out/bench_v$v.json'))
    print('variant $v', {k: d[k] for k in ('value','ms_per_step','lm_iterations_per_sec','stage_ms')})
    print('  k_eval ms', d['roofline']['ms_per_launch'], 'fp64 frac', d['roofline_fp64']['frac'], 'e2e', d['e2e']['lm_iterations_per_sec'])
except Exception as e:
    print('bench parse failed', e); print(open('gpurun_out/bench_v$v.err').read()[-2000:])
PY
done
