#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "rtma or few_query" > gpurun_out/pytest_rtma.log 2>&1; tail -3 gpurun_out/pytest_rtma.log
ARGS="--k 8 --m 8 --n 67108864 --variant 4" bash scripts/gpu_ab_one.sh
ARGS="--k 3 --m 8 --n 67108864 --variant 4" bash scripts/gpu_ab_one.sh 2>&1 | head -4
ARGS="--k 16 --m 8 --n 33554432 --variant 4" bash scripts/gpu_ab_one.sh 2>&1 | head -4
ARGS="--k 8 --m 1 --n 67108864 --variant 4" bash scripts/gpu_ab_one.sh 2>&1 | head -4
B=./multicore-hw2_b200/nn_bench
for q in 1 2 4 8; do for sp in 0 48; do
  timeout 60 $B --k 3 --m 1024 --n 65536 --variant 1 --q $q --splits $sp --iters 21 --warmup 5 | grep -v device | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(f\"q=$q sp=$sp {d['ms_med']*1000:8.2f} us (best {d['ms_best']*1000:.2f}) fp32 {d['fp32_frac_maxclk']:.4f} {d['plan'][:110]}\")"
done; done
for n in 680 5440 21760 65536; do
  timeout 60 $B --k 3 --m 1024 --n $n --variant 1 --q 8 --iters 21 --warmup 5 | grep -v device | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(f\"n=$n {d['ms_med']*1000:8.2f} us (best {d['ms_best']*1000:.2f}) {d['plan'][:110]}\")"
done
