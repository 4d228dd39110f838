#!/bin/bash
# Evidence for profiles/: launch list of the bench command + one full capture of each hot kernel.
# Every ncu run follows a plain run of the same command line (&&).
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
TAG=${1:-r02}
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-all-configs --no-e2e"
$CMD > gpurun_out/plain_bench.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_bench_cfg4.csv $CMD > gpurun_out/ncu_bench.log 2>&1
tail -2 gpurun_out/ncu_bench.log | cut -c1-200
B=./multicore-hw2_b200/nn_bench
cap() { # name, kernel regex, nn_bench args
  local C="$B $3 --iters 2 --warmup 1"
  $C > gpurun_out/plain_$1.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:$2 -s 1 -c 1 -f -o gpurun_out/${TAG}_$1 $C > gpurun_out/ncu_$1.log 2>&1
  tail -1 gpurun_out/ncu_$1.log | cut -c1-150
  # gpurun brings back at most 64 MiB: keep the raw metric table of every capture, the report itself only where asked
  ncu -i gpurun_out/${TAG}_$1.ncu-rep --page raw --csv > gpurun_out/${TAG}_$1.raw.csv 2>/dev/null
  if [ "$4" != "keep" ]; then rm -f gpurun_out/${TAG}_$1.ncu-rep; fi
}
cap cfg1_qreg nn_qreg "--k 3 --m 1024 --n 65536 --fused 1" keep
cap cfg2_qreg nn_qreg "--k 16 --m 4096 --n 1048576 --fused 1"
cap cfg3_rtma nn_rtma "--k 8 --m 8 --n 67108864 --fused 1"
cap cfg4_qreg nn_qreg "--k 16 --m 65536 --n 16777216 --fused 1" keep
cap cfg5s_qreg nn_qreg "--k 3 --m 65536 --n 1048576 --fused 1"
cap m100k8_qflex nn_qflex "--k 8 --m 100 --n 4194304 --fused 1" keep
cap m100k3_qflex nn_qflex "--k 3 --m 100 --n 4194304 --fused 1"
cap cfg3shard_rtma nn_rtma "--k 8 --m 8 --n 8388608 --fused 1"
cap m1_rreg nn_rreg "--k 8 --m 1 --n 67108864 --fused 1"
cap repack_k16 nn_repack "--repack 1 --k 16 --n 16777216"
rm -f gpurun_out/plain_*.log; ls -la gpurun_out/${TAG}_*; du -sh gpurun_out
