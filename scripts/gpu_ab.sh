#!/bin/bash
# A/B of build variants: default build vs build/alt_*; plus the standard sweeps.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
for d in multicore-hw2_b200 build/alt_*; do
  [ -x $d/nn_bench ] || continue
  tag=$(basename $d)
  timeout 300 $d/nn_bench --sweep ${SWEEP:-math} --iters 5 > gpurun_out/ab_${tag}.jsonl 2>&1
  echo "== $tag"; grep nearest_keys gpurun_out/ab_${tag}.jsonl | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(f\"{d['tag']:10s} k={d['k']:2d} m={d['m']:6d} {d['ms_med']:8.4f} ms  fp32 {d['fp32_frac_maxclk']:.4f}  {d['GBps']:7.1f} GB/s  {d['plan'][:95]}\")"
done
