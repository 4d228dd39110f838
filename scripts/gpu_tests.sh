#!/bin/bash
# what the driver runs at round end on one GPU: smoke(), the GPU suite, a short bench line
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
( time python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 ) 2>&1 | tail -5
( time timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 ) 2>&1 | tail -9
( time python bench.py --steps 5 --warmup 3 > gpurun_out/bench_short.json 2> gpurun_out/bench_short.err ) 2>&1 | grep real
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_short.json").read().strip().splitlines()[-1])
print("cfg4 ms/step %.3f frac %.3f e2e %.1f pinned %.1f parity %s" % (d["ms_per_step"], d["roofline"]["frac"], d["e2e"]["ms_per_call"], d["e2e"]["pinned"]["ms_per_call"], d["parity_spot_check"]))
for k, v in (d["all_configs"] or {}).items():
    if "ms_per_step" in v:
        print(k, "ms %.4f frac %.3f e2e %.3f ms parity %s" % (v["ms_per_step"], v["roofline"]["frac"], v["e2e"]["ms_per_call"], v["parity_spot_check"]))
PY
