#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
B=./multicore-hw2_b200/nn_bench
show() { python -c "
import sys,json
for l in sys.stdin:
    if 'nearest_keys' not in l: continue
    d=json.loads(l); print(f\"$1 {d['ms_med']*1000:9.2f} us (best {d['ms_best']*1000:.2f}) fp32 {d['fp32_frac_maxclk']:.4f} mism {d['mismatch_vs_plain']} {d['plan'][:120]}\")"; }
for q in 0 2 4 8; do timeout 60 $B --k 3 --m 1024 --n 65536 --variant 1 --q $q --iters 21 --warmup 5 --check 1 | show "cfg1 q=$q"; done
timeout 60 $B --k 16 --m 4096 --n 1048576 --iters 7 --check 1 | show cfg2
timeout 60 $B --k 3 --m 65536 --n 1048576 --iters 5 --check 1 | show cfg5/16
timeout 60 $B --k 16 --m 65536 --n 262144 --iters 5 | show cfg4/64
C="$B --k 3 --m 1024 --n 65536 --iters 2 --warmup 1 --q 4"
$C > gpurun_out/plain_cfg1.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:nn_qreg -s 1 -c 1 -f -o gpurun_out/r01_cfg1_auto $C > gpurun_out/ncu_cfg1.log 2>&1
tail -1 gpurun_out/ncu_cfg1.log
