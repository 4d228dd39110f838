#!/bin/bash
cd "$(dirname "$0")/.."
for d in build/alt_order0 multicore-hw2_b200 build/alt_order2; do
 for g in 1 128; do
  for a in "--k 16 --m 65536 --n 16777216 --iters 3" "--k 16 --m 4096 --n 1048576 --iters 9" "--k 8 --m 65536 --n 4194304 --iters 5" "--k 3 --m 1024 --n 65536 --iters 31" "--k 8 --m 100 --n 4194304 --iters 9"; do
    $d/nn_bench $a --waves 8 --fused 1 --warmup 2 --check 1 --opt qgroup=$g 2>/dev/null | grep -v '"device"' | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('$(basename $d) qgroup=$g'.ljust(30), f\"k={d['k']:2d} m={d['m']:7d} n={d['n']:9d} {d['ms_med']*1e3:10.1f} us best {d['ms_best']*1e3:10.1f} fp32 {d['fp32_frac_maxclk']:.4f} mism {d['mismatch_vs_plain']}\")"
  done
 done
done
