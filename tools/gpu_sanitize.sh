#!/bin/bash
# compute-sanitizer over every kernel of the iteration graph at small sizes.
mkdir -p gpurun_out
TAG=${TAG:-r02}
for tool in memcheck racecheck synccheck initcheck; do
  for c in ${CASES:-dense pairs mono robust}; do
    log=gpurun_out/${TAG}_sanitizer_${tool}_${c}.log
    timeout ${SAN_TIMEOUT:-420} compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_case.py $c > $log 2>&1
    echo "rc=$?" >> $log
    echo "== $tool $c: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|rc=' $log | tr '\n' ' ')"
  done
done
