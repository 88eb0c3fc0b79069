#!/bin/bash
# A/B of the config-4 Schur kernels: per-kernel launch times (ncu launch list) for several builds.
mkdir -p gpurun_out
summ() { python - "$1" <<'PY'
import csv, collections, sys
rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
h = rows[0]; ki = h.index('Kernel Name'); vi = h.index('Metric Value')
d = collections.defaultdict(list)
for r in rows[1:]:
    d[r[ki].split('(')[0]].append(float(r[vi].replace(',', '')))
print(sys.argv[1], {k: round(sum(v)/len(v)/1e3, 1) for k, v in d.items()})
PY
}
for lib in ${LIBS:-default chunk32 minb1 sf3}; do
  if [ "$lib" = default ]; then unset TSCM_LIB_PATH; else export TSCM_LIB_PATH=$PWD/build_ab/libtscm_$lib.so; fi
  timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_schur_frames|k_pair_blocks|k_schur_pairs|k_reduce_pairs' -c 30 --csv --log-file gpurun_out/ab_$lib.csv python tools/stress_cfg4.py --frames ${FRAMES:-40000} --timed-iterations 3 --out gpurun_out/stress_ab_$lib.json > gpurun_out/ab_$lib.log 2>&1
  summ gpurun_out/ab_$lib.csv
done
unset TSCM_LIB_PATH
if [ -n "$FULL" ]; then
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_schur_pairs -s 3 -c 1 -f -o gpurun_out/prof_cfg4_k_schur_pairs2 python tools/stress_cfg4.py --frames ${FRAMES:-40000} --timed-iterations 3 --out gpurun_out/stress_cfg4_ncu.json > gpurun_out/ncu_cfg4_pairs2.log 2>&1
fi
