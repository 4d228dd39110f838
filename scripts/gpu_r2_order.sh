#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
for d in multicore-hw2_b200 build/alt_splitfast; do
  for a in "--k 16 --m 65536 --n 16777216 --waves 8 --iters 3" "--k 16 --m 4096 --n 1048576 --waves 8 --iters 9" "--k 3 --m 1024 --n 65536 --waves 8 --iters 31" "--k 3 --m 1048576 --n 1048576 --waves 8 --iters 3" "--k 8 --m 65536 --n 4194304 --waves 8 --iters 3" "--k 16 --m 1024 --n 1048576 --waves 8 --iters 9"; do
    $d/nn_bench $a --fused 1 --warmup 2 | grep -v '"device"' | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('$(basename $d)'.ljust(20), f\"k={d['k']:2d} m={d['m']:7d} n={d['n']:9d} {d['ms_med']*1e3:10.1f} us best {d['ms_best']*1e3:10.1f} fp32 {d['fp32_frac_maxclk']:.4f} | {d['plan'][60:130]}\")"
  done
  C="$d/nn_bench --k 16 --m 65536 --n 16777216 --waves 8 --fused 1 --iters 1 --warmup 1"
  ncu --metrics dram__bytes_read.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:nn_qreg -s 1 -c 1 --csv $C 2>/dev/null | grep -E "dram__bytes_read|gpu__time|hit_rate" | cut -d, -f9,13,14,15
done
