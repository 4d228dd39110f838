#!/bin/bash
# First run of the reference-stream (TMA ring) kernel: parity, then cfg3 timing against kernel B.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "rtma or few_query" > gpurun_out/pytest_rtma.log 2>&1; tail -5 gpurun_out/pytest_rtma.log
B=./multicore-hw2_b200/nn_bench
for v in 2 4; do
  for m in 8 1 2 4 16; do
    timeout 120 $B --k 8 --m $m --n 67108864 --variant $v --iters 7 --check 1 | grep -v device | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(f\"v$v m=$m {d['ms_med']:8.4f} ms (best {d['ms_best']:.4f}) fp32 {d['fp32_frac_maxclk']:.4f}  {d['GBps']:7.1f} GB/s mism {d['mismatch_vs_plain']} {d['plan'][:90]}\")"
  done
done
for k in 3 16; do for v in 2 4; do
    timeout 120 $B --k $k --m 8 --n 33554432 --variant $v --iters 5 --check 1 | grep -v device | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(f\"v$v k=$k m=8 {d['ms_med']:8.4f} ms (best {d['ms_best']:.4f}) fp32 {d['fp32_frac_maxclk']:.4f}  {d['GBps']:7.1f} GB/s mism {d['mismatch_vs_plain']} {d['plan'][:90]}\")"
done; done
