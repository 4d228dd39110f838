#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
B=./multicore-hw2_b200/nn_bench
for mth in 0 1 2; do
C="$B --k 16 --m 4096 --n 1048576 --iters 2 --warmup 1 --variant 1 --q 4 --math $mth"
$C > gpurun_out/plain_m$mth.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:nn_qreg -s 1 -c 1 -f -o gpurun_out/prof_k16_math$mth $C > gpurun_out/ncu_m$mth.log 2>&1
done
ls -la gpurun_out/*.ncu-rep
