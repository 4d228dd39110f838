#!/usr/bin/env python
"""nn_bench --sweep cross output (JSON lines) -> profiles/r02_fewquery_crossover.json."""
import json
import sys

rows = {}
for l in open(sys.argv[1]):
    if not l.startswith("{") or '"op":"nearest_keys"' not in l:
        continue
    d = json.loads(l)
    key = (d["k"], d["m"], d["n"])
    r = rows.setdefault(key, {"k": d["k"], "m": d["m"], "n": d["n"]})
    r[d["tag"] + "_us"] = round(d["ms_med"] * 1e3, 2)
    if d["tag"] == "auto":
        r["auto_plan"] = d["plan"].split()[0]
out = {"source": "multicore-hw2_b200/nn_bench --sweep cross --iters 7 on one B200 (kernel only, CUDA events, median of 7)",
       "rows": [rows[k] for k in sorted(rows)]}
json.dump(out, open(sys.argv[2], "w"), indent=0)
print(len(out["rows"]), "shapes")
