#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "rtma or few_query" > gpurun_out/pytest_rtma.log 2>&1; tail -2 gpurun_out/pytest_rtma.log
B=./multicore-hw2_b200/nn_bench
show() { python -c "
import sys,json
for l in sys.stdin:
    if 'nearest_keys' not in l: continue
    d=json.loads(l); print(f\"$1 {d['ms_med']*1000:9.2f} us (best {d['ms_best']*1000:.2f}) fp32 {d['fp32_frac_maxclk']:.4f} {d['GBps']:7.1f} GB/s mism {d['mismatch_vs_plain']} {d['plan'][:80]}\")"; }
for k in 3 5 8 12 13 16; do for v in 2 4; do timeout 60 $B --k $k --m 8 --n 16777216 --variant $v --iters 7 --check 1 | show "k=$k m=8 v=$v"; done; done
for m in 5 6 16 24 32 48; do for v in 2 4 1; do timeout 60 $B --k 8 --m $m --n 16777216 --variant $v --iters 7 | show "k=8 m=$m v=$v"; done; done
for m in 64 96 128; do for v in 4 1; do timeout 60 $B --k 8 --m $m --n 16777216 --variant $v --iters 7 | show "k=8 m=$m v=$v"; done; done
timeout 60 $B --k 8 --m 8 --n 67108864 --variant 4 --iters 9 | show "cfg3 v=4"
