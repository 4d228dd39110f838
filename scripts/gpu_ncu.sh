#!/bin/bash
# ncu captures of the two hot kernels (one launch each) + the FP32 issue probes.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
B=./multicore-hw2_b200/nn_bench
$B --sweep probe > gpurun_out/probe.jsonl 2>&1
cat gpurun_out/probe.jsonl
C2="$B --k 16 --m 4096 --n 1048576 --iters 2 --warmup 1"
C3="$B --k 8 --m 8 --n 67108864 --iters 2 --warmup 1"
$C2 > gpurun_out/plain_cfg2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:nn_qreg -s 1 -c 1 -f -o gpurun_out/prof_cfg2 $C2 > gpurun_out/ncu_cfg2.log 2>&1
$C3 > gpurun_out/plain_cfg3.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:nn_rreg -s 1 -c 1 -f -o gpurun_out/prof_cfg3 $C3 > gpurun_out/ncu_cfg3.log 2>&1
tail -3 gpurun_out/ncu_cfg2.log gpurun_out/ncu_cfg3.log
ls -la gpurun_out
