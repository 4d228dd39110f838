#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
B=./multicore-hw2_b200/nn_bench
show() { python -c "
import sys,json
for l in sys.stdin:
    if 'nearest_keys' not in l: continue
    d=json.loads(l); print(f\"$1 {d['ms_med']*1000:9.2f} us (best {d['ms_best']*1000:.2f}) fp32 {d['fp32_frac_maxclk']:.4f} mism {d['mismatch_vs_plain']} {d['plan'][:120]}\")"; }
for q in 0 1 2 4 8; do timeout 60 $B --k 3 --m 1024 --n 65536 --variant 1 --q $q --iters 21 --warmup 5 --check 1 | show "cfg1 q=$q"; done
for sp in 74 148 222 296 444 592; do timeout 60 $B --k 3 --m 1024 --n 65536 --variant 1 --q 4 --splits $sp --iters 21 --warmup 5 | show "cfg1 q=4 sp=$sp"; done
for sp in 74 148 296 592; do timeout 60 $B --k 3 --m 1024 --n 65536 --variant 1 --q 8 --splits $sp --iters 21 --warmup 5 | show "cfg1 q=8 sp=$sp"; done
for sp in 37 74 148 296; do timeout 60 $B --k 3 --m 1024 --n 65536 --variant 1 --q 2 --splits $sp --iters 21 --warmup 5 | show "cfg1 q=2 sp=$sp"; done
timeout 60 $B --k 16 --m 4096 --n 1048576 --iters 7 --check 1 | show cfg2
timeout 60 $B --k 3 --m 65536 --n 1048576 --iters 5 --check 1 | show cfg5/16
timeout 60 $B --k 16 --m 65536 --n 262144 --iters 5 | show cfg4/64
timeout 60 $B --k 8 --m 8 --n 67108864 --iters 7 --variant 4 | show cfg3-rtma
