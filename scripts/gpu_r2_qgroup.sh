#!/bin/bash
cd "$(dirname "$0")/.."
B=./multicore-hw2_b200/nn_bench
for g in 1 8 32 64 128; do
  for a in "--k 16 --m 65536 --n 16777216 --iters 3" "--k 8 --m 65536 --n 4194304 --iters 5" "--k 16 --m 4096 --n 1048576 --iters 9" "--k 3 --m 1048576 --n 1048576 --iters 3"; do
    $B $a --waves 8 --fused 1 --warmup 2 --opt qgroup=$g 2>/dev/null | grep -v '"device"' | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('qgroup=$g'.ljust(12), f\"k={d['k']:2d} m={d['m']:7d} n={d['n']:9d} {d['ms_med']*1e3:10.1f} us best {d['ms_best']*1e3:10.1f} fp32 {d['fp32_frac_maxclk']:.4f}\")"
  done
  C="$B --k 16 --m 65536 --n 16777216 --waves 8 --fused 1 --iters 1 --warmup 1 --opt qgroup=$g"
  ncu --metrics dram__bytes_read.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:nn_qreg -s 1 -c 1 --csv $C 2>/dev/null | grep -E "dram__bytes_read|hit_rate" | awk -F'","' '{print "   qgroup='$g'", $(NF-2), $(NF-1), $NF}'
done
