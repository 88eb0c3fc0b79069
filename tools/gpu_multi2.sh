#!/bin/bash
# Multi-GPU visit (short): weak-scaling bench on N ranks + sharded-vs-single parity.
N=${N:-2}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/smi_L.txt 2>&1
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 30 --warmup 5 > gpurun_out/bench_n${N}_weak.json 2> gpurun_out/bench_n${N}_weak.err
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 tools/dist_parity.py > gpurun_out/dist_parity_n${N}.log 2>&1
python - gpurun_out/bench_n${N}_weak.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], {k: d.get(k) for k in ('n_gpus','scaling','value','ms_per_step','lm_iterations_per_sec')}, 'e2e', d.get('e2e',{}).get('lm_iterations_per_sec'))
except Exception as e:
    print(sys.argv[1], 'FAILED', e)
PY
tail -3 gpurun_out/bench_n${N}_weak.err; tail -4 gpurun_out/dist_parity_n${N}.log
