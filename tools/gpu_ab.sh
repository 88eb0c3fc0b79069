#!/bin/bash
# A/B: bench the default library and every build_ab/*.so (TSCM_LIB_PATH)
mkdir -p gpurun_out
for lib in default build_ab/*.so; do
  if [ "$lib" = default ]; then unset TSCM_LIB_PATH; else export TSCM_LIB_PATH=$PWD/$lib; fi
  timeout 300 python bench.py --no-cpu-baseline --steps 40 > gpurun_out/ab.json 2> gpurun_out/ab.err
  python - "$lib" <<'PY'
import json, sys
try:
    d = json.load(open('gpurun_out/ab.json'))
    print(f"{sys.argv[1]:24s} it/s {d['lm_iterations_per_sec']:8.1f}  k_eval ms {d['roofline']['ms_per_launch']:.4f}  stages {d['stage_ms']}")
except Exception as e:
    print(sys.argv[1], 'failed', e, open('gpurun_out/ab.err').read()[-500:])
PY
done
