#!/bin/bash
# Kernel sweeps + parity tests.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
B=./multicore-hw2_b200/nn_bench
timeout 120 $B --sweep check > gpurun_out/sweep_check.jsonl 2> gpurun_out/sweep_check.err
python - <<'PY'
import json
bad=n=0
for l in open('gpurun_out/sweep_check.jsonl'):
    d=json.loads(l)
    if 'mismatch_vs_plain' in d: n+=1; bad+=d['mismatch_vs_plain']!=0
print('check sweep:',n,'configs, bad',bad)
PY
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
tail -4 gpurun_out/pytest_gpu.log
for s in ${SWEEPS:-math cfgs small}; do timeout 300 $B --sweep $s --iters 5 > gpurun_out/sweep_$s.jsonl 2>&1; done
cat gpurun_out/sweep_*.jsonl | grep nearest_keys | cut -c1-360
