#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
B=./multicore-hw2_b200/nn_bench
for q in 1 2 4 8; do for sp in 0 12 24 37 48 74 96; do
  timeout 60 $B --k 3 --m 1024 --n 65536 --variant 1 --q $q --splits $sp --iters 21 --warmup 5 | grep -v device | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(f\"q=$q sp=$sp {d['ms_med']*1000:8.2f} us (best {d['ms_best']*1000:.2f}) fp32 {d['fp32_frac_maxclk']:.4f} {d['plan'][:110]}\")"
done; done
