#!/bin/bash
# round 2, final single-GPU pass: full GPU suite, smoke, the driver's bench command, family crossover with the final kernels
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
( time timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 ) 2>&1 | tail -9
( time python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 ) 2>&1 | tail -5
( time python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err ) 2>&1 | grep real
tail -c 300 gpurun_out/r02_bench_n1.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_bench_n1.json").read().strip().splitlines()[-1])
print("cfg4 value %.4e ms/step %.3f frac %.3f e2e %.1f pinned %.1f index %.1f parity %s launches %d" % (
    d["value"], d["ms_per_step"], d["roofline"]["frac"], d["e2e"]["ms_per_call"], d["e2e"]["pinned"]["ms_per_call"],
    d["e2e"]["resident_index"]["ms_per_call"], d["parity_spot_check"], d["gpu_launches"]))
for k, v in (d["all_configs"] or {}).items():
    if "ms_per_step" in v:
        print(k, "ms %.4f frac %.3f e2e %.3f ms parity %s" % (v["ms_per_step"], v["roofline"]["frac"], v["e2e"]["ms_per_call"], v["parity_spot_check"]))
PY
B=./multicore-hw2_b200/nn_bench
$B --sweep cross --iters 7 --warmup 2 2> gpurun_out/r2_cross.err | grep -v '"device"' > gpurun_out/r2_cross.jsonl
python scripts/make_crossover.py gpurun_out/r2_cross.jsonl gpurun_out/r02_fewquery_crossover.json
