#!/bin/bash
# Phase timeline of the query-register kernel on small searches (debug build: make EXTRA=-DNN_QREG_TIMELINE
# OUT=build/alt_tl); prints where the CTAs' time goes.
cd "$(dirname "$0")/.."
B=build/alt_tl/nn_bench
mkdir -p gpurun_out
{
for q in 0 4 8; do
  $B --k 3 --m 1024 --n 65536 --fused 1 --timeline 1 --q $q --iters 50 --warmup 10
done
$B --k 16 --m 1024 --n 65536 --fused 1 --timeline 1 --iters 50 --warmup 10
./multicore-hw2_b200/nn_bench --k 3 --m 1024 --n 65536 --fused 1 --iters 200 --warmup 20
./multicore-hw2_b200/nn_bench --k 3 --m 1024 --n 64 --fused 1 --iters 200 --warmup 20
} > gpurun_out/r2_timeline.txt 2>&1
cat gpurun_out/r2_timeline.txt
