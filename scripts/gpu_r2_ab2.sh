#!/bin/bash
cd "$(dirname "$0")/.."
for d in multicore-hw2_b200 build/alt_q4u1 build/alt_q4u1b5 build/alt_b5; do
  for a in "--k 3 --m 100" "--k 4 --m 100" "--k 5 --m 100" "--k 8 --m 100" "--k 12 --m 100" "--k 16 --m 100" "--k 3 --m 64" "--k 8 --m 64" "--k 3 --m 200" "--k 8 --m 200" "--k 3 --m 32" "--k 8 --m 32"; do
    $d/nn_bench $a --n 4194304 --variant 5 --iters 9 2>/dev/null | grep -v '"device"' | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('$(basename $d)'.ljust(20), f\"k={d['k']:2d} m={d['m']:4d} {d['ms_med']*1e3:8.1f} us fp32 {d['fp32_frac_maxclk']:.3f} | {d['plan'][:100]}\")"
  done
done
