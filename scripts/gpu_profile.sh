#!/bin/bash
# Evidence for profiles/: launch list of the bench command + one full capture of each hot kernel.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
CMD="python bench.py --steps 2 --warmup 3"
$CMD > gpurun_out/plain_bench.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench.csv $CMD > gpurun_out/ncu_bench.log 2>&1
tail -2 gpurun_out/ncu_bench.log | cut -c1-200
B=./multicore-hw2_b200/nn_bench
C2="$B --k 16 --m 4096 --n 1048576 --iters 2 --warmup 1"
$C2 > gpurun_out/plain_cfg2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:nn_qreg -s 1 -c 1 -f -o gpurun_out/r01_cfg2_qreg $C2 > gpurun_out/ncu_cfg2.log 2>&1
C3="$B --k 8 --m 8 --n 67108864 --iters 2 --warmup 1"
$C3 > gpurun_out/plain_cfg3.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:nn_rreg -s 1 -c 1 -f -o gpurun_out/r01_cfg3_rreg $C3 > gpurun_out/ncu_cfg3.log 2>&1
C5="$B --k 3 --m 65536 --n 1048576 --iters 2 --warmup 1"
$C5 > gpurun_out/plain_cfg5.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:nn_qreg -s 1 -c 1 -f -o gpurun_out/r01_cfg5s_qreg $C5 > gpurun_out/ncu_cfg5.log 2>&1
CR="$B --repack 1 --k 16 --n 16777216 --iters 2 --warmup 1"
$CR > gpurun_out/plain_repack.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:nn_repack -s 1 -c 1 -f -o gpurun_out/r01_repack_k16 $CR > gpurun_out/ncu_repack.log 2>&1
ls -la gpurun_out/*.ncu-rep gpurun_out/launches_bench.csv
