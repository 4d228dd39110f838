#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
B=./multicore-hw2_b200/nn_bench
$B --sweep cfgs --iters 5 > gpurun_out/sweep_cfgs.jsonl 2>&1
for ldg in 0 1; do
C="$B --k 8 --m 8 --n 67108864 --iters 2 --warmup 1 --ldg $ldg"
$C > gpurun_out/plain_c3_$ldg.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"nn_r" -s 1 -c 1 -f -o gpurun_out/prof_cfg3_ldg$ldg $C > gpurun_out/ncu_c3_$ldg.log 2>&1
done
ls -la gpurun_out/*.ncu-rep
