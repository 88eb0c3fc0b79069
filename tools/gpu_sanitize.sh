#!/bin/bash
# compute-sanitizer over every kernel of the iteration graph at small sizes.
# CASES / TOOLS select; with >= 2 GPUs also the single-process group and a 2-rank torchrun solve.
mkdir -p gpurun_out
TAG=${TAG:-r02}
for tool in ${TOOLS:-memcheck racecheck synccheck initcheck}; do
  for c in ${CASES:-dense pairs mono robust masks init}; do
    log=gpurun_out/${TAG}_sanitizer_${tool}_${c}.log
    timeout ${SAN_TIMEOUT:-420} compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_case.py $c > $log 2>&1
    echo "rc=$?" >> $log
    echo "== $tool $c: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|rc=' $log | tr '\n' ' ')"
  done
done
if [ "$(nvidia-smi -L | wc -l)" -ge 2 ]; then
  for tool in memcheck synccheck; do
    log=gpurun_out/${TAG}_sanitizer_${tool}_group2.log
    timeout ${SAN_TIMEOUT:-420} compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_case.py group > $log 2>&1
    echo "rc=$?" >> $log
    echo "== $tool group2: $(grep -E 'ERROR SUMMARY|rc=' $log | tr '\n' ' ')"
  done
  log=gpurun_out/${TAG}_sanitizer_memcheck_ranks2.log
  DIST_CFGS=2 timeout ${SAN_TIMEOUT:-420} compute-sanitizer --tool memcheck --target-processes all --print-limit 20 \
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29577 tools/dist_parity.py > $log 2>&1
  echo "rc=$?" >> $log
  echo "== memcheck ranks2: $(grep -E 'ERROR SUMMARY|DIST PARITY|rc=' $log | tr '\n' ' ')"
fi
