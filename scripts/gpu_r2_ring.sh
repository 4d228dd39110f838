#!/bin/bash
cd "$(dirname "$0")/.."
( timeout 600 python -m pytest tests -m gpu -x -q -k "phased or fixture or one_launch or cross_check" 2>&1 | tail -4 )
B=./multicore-hw2_b200/nn_bench
for deep in 0 1; do for k in 3 8 16; do for m in 16 25 32 48 64 100 128; do
  $B --k $k --m $m --n 4194304 --variant 5 --iters 7 --check 1 --opt flex_deep_ring=$deep 2>/dev/null | grep -v '"device"' | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('deep=$deep', f\"k={d['k']:2d} m={d['m']:4d} {d['ms_med']*1e3:8.1f} us fp32 {d['fp32_frac_maxclk']:.3f} mism {d['mismatch_vs_plain']} | {d['plan'][:110]}\")"
done; done; done
for deep in 0 1; do for k in 3 8; do for m in 32 64; do
  $B --k $k --m $m --n 1048576 --variant 5 --iters 9 --opt flex_deep_ring=$deep 2>/dev/null | grep -v '"device"' | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('deep=$deep', f\"k={d['k']:2d} m={d['m']:4d} n=2^20 {d['ms_med']*1e3:8.1f} us fp32 {d['fp32_frac_maxclk']:.3f} | {d['plan'][:110]}\")"
done; done; done
