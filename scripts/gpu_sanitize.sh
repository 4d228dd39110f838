#!/bin/bash
# memcheck of every kernel family on small tie-heavy cases (one compute-sanitizer tool per call).
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
TOOL=${1:-memcheck}
B=./multicore-hw2_b200/nn_bench
$B --sweep check --iters 1 > gpurun_out/plain_check.log 2>&1 || { tail -5 gpurun_out/plain_check.log; exit 1; }
grep -c '"mismatch_vs_plain":0' gpurun_out/plain_check.log; grep -c nearest_keys gpurun_out/plain_check.log
timeout 1500 compute-sanitizer --tool $TOOL --print-limit 20 $B --sweep check --iters 1 > gpurun_out/sanitize_$TOOL.log 2>&1
tail -4 gpurun_out/sanitize_$TOOL.log
