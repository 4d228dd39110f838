#!/bin/bash
# Quick state check: smoke, GPU tests, the headline bench and one extra workload ($1, default cfg4).
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
W=${1:-cfg4}
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err; cut -c1-600 gpurun_out/bench.json
timeout 900 python bench.py --workload $W --steps 3 > gpurun_out/bench_$W.json 2> gpurun_out/bench_$W.err; tail -2 gpurun_out/bench_$W.err; cut -c1-900 gpurun_out/bench_$W.json
