#!/bin/bash
# A/B of the query-register kernel's super-chunk loop (option qreg_super = chunks per (best, where) update),
# every k at m = 65536, n = 2^21, plus the BASELINE shapes.  Output: gpurun_out/r2_super.txt
cd "$(dirname "$0")/.."
B=./multicore-hw2_b200/nn_bench
mkdir -p gpurun_out
{
for k in 3 4 5 6 7 8 9 10 11 12 13 14 15 16; do
  for sc in 1 8 16 64; do
    $B --k $k --m 65536 --n 2097152 --fused 1 --iters 2 --warmup 1 --opt qreg_super=$sc --tag k${k}_sc$sc
  done
done
for sc in 1 32; do
  $B --k 16 --m 65536 --n 16777216 --fused 1 --iters 2 --warmup 1 --opt qreg_super=$sc --tag cfg4_sc$sc
done
$B --k 3 --m 1048576 --n 1048576 --fused 1 --iters 3 --warmup 1 --opt qreg_super=1 --tag cfg5_sc1
$B --k 16 --m 4096 --n 1048576 --fused 1 --iters 20 --warmup 3 --opt qreg_super=1 --tag cfg2_sc1
$B --k 3 --m 1024 --n 65536 --fused 1 --iters 200 --warmup 20 --opt qreg_super=1 --tag cfg1_sc1
} > gpurun_out/r2_super.txt 2>&1
grep -v '"device"' gpurun_out/r2_super.txt | python3 -c "
import sys,json
for l in sys.stdin:
    l=l.strip()
    if l.startswith('{'):
        r=json.loads(l); print(r['tag'], r['ms_med'], r['ms_best'], 'frac %.4f'%r['fp32_frac_maxclk'], r['plan'][:44])
    else: print(l)
"
