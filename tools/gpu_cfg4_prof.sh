#!/bin/bash
# Config-4 Schur kernels: launch times (ncu, serialised) and one full capture each.
mkdir -p gpurun_out
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_schur_frames|k_pair_blocks|k_schur_pairs|k_reduce_pairs|k_reduce_s|k_backsub|k_solve|k_eval5|k_view_blocks|k_post_eval' -c 120 --csv --log-file gpurun_out/launches_cfg4.csv python tools/stress_cfg4.py --frames ${FRAMES:-40000} --timed-iterations 3 --out gpurun_out/stress_cfg4_ncu.json > gpurun_out/ncu_cfg4.log 2>&1
for k in k_schur_pairs k_schur_frames; do
  timeout 200 ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -f -o gpurun_out/prof_cfg4_$k python tools/stress_cfg4.py --frames ${FRAMES:-40000} --timed-iterations 3 --out gpurun_out/stress_cfg4_ncu.json > gpurun_out/ncu_cfg4_$k.log 2>&1
done
python - <<'PY'
import csv, collections
rows = list(csv.reader(l for l in open('gpurun_out/launches_cfg4.csv') if l.startswith('"')))
h = rows[0]; ki = h.index('Kernel Name'); vi = h.index('Metric Value')
d = collections.defaultdict(list)
for r in rows[1:]:
    d[r[ki].split('(')[0]].append(float(r[vi].replace(',', '')))
for k, v in d.items():
    print(f"{k:40s} n={len(v):3d} mean {sum(v)/len(v)/1e3:9.1f} us  min {min(v)/1e3:9.1f}")
PY
