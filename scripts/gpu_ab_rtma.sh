#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
ARGS="--k 8 --m 8 --n 67108864 --variant 4" bash scripts/gpu_ab_one.sh
B=./multicore-hw2_b200/nn_bench
C3="$B --k 8 --m 8 --n 67108864 --variant 4 --iters 2 --warmup 1"
$C3 > gpurun_out/plain_rtma.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:nn_rtma -s 1 -c 1 -f -o gpurun_out/r01_cfg3_rtma $C3 > gpurun_out/ncu_rtma.log 2>&1
tail -2 gpurun_out/ncu_rtma.log
