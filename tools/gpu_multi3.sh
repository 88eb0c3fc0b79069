#!/bin/bash
# Multi-GPU A/B: peer-memory exchange (default) vs NCCL (TSCM_P2P=0): parity first, then weak bench.
N=${N:-2}
mkdir -p gpurun_out
for mode in ${MODES:-1 0}; do
  export TSCM_P2P=$mode
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$mode tools/dist_parity.py > gpurun_out/dist_parity_n${N}_p2p$mode.log 2>&1
  tail -3 gpurun_out/dist_parity_n${N}_p2p$mode.log
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2952$mode bench.py --gpus $N --steps 30 --warmup 5 > gpurun_out/bench_n${N}_weak_p2p$mode.json 2> gpurun_out/bench_n${N}_weak_p2p$mode.err
  python - gpurun_out/bench_n${N}_weak_p2p$mode.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], {k: d.get(k) for k in ('n_gpus','scaling','value','ms_per_step','lm_iterations_per_sec')}, 'e2e', d.get('e2e',{}).get('lm_iterations_per_sec'))
except Exception as e:
    print(sys.argv[1], 'FAILED', e); print(open(sys.argv[1].replace('.json','.err')).read()[-1500:])
PY
done
