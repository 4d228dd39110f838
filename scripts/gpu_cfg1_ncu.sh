#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
B=./multicore-hw2_b200/nn_bench
rm -f gpurun_out/cfg1_launches.txt
for q in 1 2 8; do
  C="$B --k 3 --m 1024 --n 65536 --variant 1 --q $q --iters 3 --warmup 2"
  $C > gpurun_out/plain_cfg1.log 2>&1 && ncu --metrics gpu__time_duration.sum,sm__cycles_elapsed.max,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum --clock-control none -k regex:nn_qreg --csv $C 2>/dev/null | grep -E "gpu__time_duration|sm__cycles_elapsed|pipe_fma_cycles|inst_executed" | awk -F'","' -v q=$q '{print "q="q, $5, $(NF-2), $NF}' >> gpurun_out/cfg1_launches.txt
done
C="$B --k 3 --m 1024 --n 680 --variant 1 --q 8 --iters 3 --warmup 2"
$C > gpurun_out/plain_cfg1.log 2>&1 && ncu --metrics gpu__time_duration.sum,sm__cycles_elapsed.max --clock-control none -k regex:nn_qreg --csv $C 2>/dev/null | grep -E "gpu__time_duration|sm__cycles_elapsed" | awk -F'","' '{print "n=680 q=8", $5, $(NF-2), $NF}' >> gpurun_out/cfg1_launches.txt
cat gpurun_out/cfg1_launches.txt
C="$B --k 3 --m 1024 --n 65536 --variant 1 --q 2 --iters 2 --warmup 1"
$C > gpurun_out/plain_cfg1.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:nn_qreg -s 1 -c 1 -f -o gpurun_out/r01_cfg1_q2 $C > gpurun_out/ncu_cfg1.log 2>&1
tail -1 gpurun_out/ncu_cfg1.log
