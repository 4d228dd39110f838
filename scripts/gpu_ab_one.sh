#!/bin/bash
# A/B of build variants on one nn_bench command line: ARGS="--k 8 --m 8 --n 67108864"
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
for d in multicore-hw2_b200 build/alt_*; do
  [ -x $d/nn_bench ] || continue
  echo "== $(basename $d)"; $d/nn_bench $ARGS --iters 7 | grep -v device | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(f\"{d['ms_med']:8.4f} ms (best {d['ms_best']:.4f}) fp32 {d['fp32_frac_maxclk']:.4f}  {d['GBps']:7.1f} GB/s  {d['plan'][:100]}\")"
done
