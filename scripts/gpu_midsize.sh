#!/bin/bash
# Mid-size shapes (the reference's README ids 7, 10, 11 and neighbours): auto plan vs forced tile widths.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
B=./multicore-hw2_b200/nn_bench
QS=${QS:-"0 8 4 2 1"}
show() { python -c "
import sys,json
for l in sys.stdin:
    if 'nearest_keys' not in l: continue
    d=json.loads(l); print(f\"$1 {d['ms_med']*1000:9.1f} us fp32 {d['fp32_frac_maxclk']:.3f} mism {d['mismatch_vs_plain']} {d['plan'][:112]}\")"; }
for shape in "3 1024 65536" "16 1024 65536" "3 1024 1048576" "16 1024 1048576" "8 512 262144" "16 256 1048576" "5 2000 500000" "12 200 4000000"; do
  set -- $shape
  for q in $QS; do
    timeout 60 $B --k $1 --m $2 --n $3 --variant 1 --q $q --iters 9 --check ${CHECK:-0} 2>/dev/null | show "k=$1 m=$2 n=$3 q=$q"
  done
done
