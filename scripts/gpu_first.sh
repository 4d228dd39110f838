#!/bin/bash
# First GPU visit: parity tests, kernel sweeps, contract bench.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
{ nproc; lscpu | grep -E "Model name|Socket|Core|Thread" ; free -g | head -2; nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit,memory.total --format=csv; } > gpurun_out/box.txt 2>&1
export LD_LIBRARY_PATH=$PWD/multicore-hw2_b200:$LD_LIBRARY_PATH
timeout 120 ./multicore-hw2_b200/nn_bench --sweep check > gpurun_out/sweep_check.jsonl 2> gpurun_out/sweep_check.err
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
tail -5 gpurun_out/pytest_gpu.log
timeout 300 ./multicore-hw2_b200/nn_bench --sweep math --iters 5 > gpurun_out/sweep_math.jsonl 2>&1
timeout 300 ./multicore-hw2_b200/nn_bench --sweep cfgs --iters 5 > gpurun_out/sweep_cfgs.jsonl 2>&1
timeout 300 ./multicore-hw2_b200/nn_bench --sweep small --iters 5 > gpurun_out/sweep_small.jsonl 2>&1
timeout 120 ./multicore-hw2_b200/nn_bench --sweep repack --iters 5 > gpurun_out/sweep_repack.jsonl 2>&1
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json
cat gpurun_out/sweep_math.jsonl gpurun_out/sweep_cfgs.jsonl | cut -c1-330
