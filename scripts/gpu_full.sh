#!/bin/bash
# Round-end style check: smoke, GPU tests, both bench arms.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cat gpurun_out/bench_ref.json | cut -c1-400
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json
for w in cfg1 cfg3 cfg5; do timeout 900 python bench.py --workload $w --steps 5 > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; tail -2 gpurun_out/bench_$w.err; cat gpurun_out/bench_$w.json; done
