#!/bin/bash
# All five BASELINE configs through bench.py on one GPU (+ the reference arm on the headline config).
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cut -c1-300 gpurun_out/bench_ref.json
for w in cfg2 cfg1 cfg3 cfg5 cfg4; do
  st=5; [ $w = cfg4 ] && st=3; [ $w = cfg2 ] && st=10
  timeout 1200 python bench.py --workload $w --steps $st > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; tail -2 gpurun_out/bench_$w.err
  python - <<PY
import json
for l in open("gpurun_out/bench_$w.json"):
    if l.startswith("{"):
        d=json.loads(l); r=d["roofline"]; e=d["e2e"]; c=d["cpu_baseline"]
        print("$w value %.4e pairs/s  step %.4f ms  kernel %.4f ms  %s frac %.3f (fp32 %.3f hbm %.3f)  e2e %.4e (%.3f ms)  cpu %.3e x%d  clocks %s  spot %s" % (d["value"], d["ms_per_step"], r["kernel_ms"], r["bound"], r["frac"], r["fp32"]["frac"], r["hbm"]["frac"], e["value"], e["ms_per_call"], c["value"], c["cores"], d["clocks"]["sm_mhz"], c.get("parity_spot_check")))
        print("   plan:", d["config"]["plan"])
PY
done
cp gpurun_out/bench_cfg2.json gpurun_out/bench.json
