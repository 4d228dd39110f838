#!/usr/bin/env python
"""Wall clock of cudaCallback (malloc'ed inputs) on 1, 2, 4, ... GPUs for the TA samples and the BASELINE
configs: the data nn_b200_plan_gpus is fitted to / checked against (profiles/r02_gpu_count.json)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import multicore_hw2_b200 as nn  # noqa: E402
import torch  # noqa: E402

vis = torch.cuda.device_count()
counts = [g for g in (1, 2, 4, 8) if g <= vis]
shapes = [("ta0", 3, 1, 2), ("ta1", 3, 2, 8), ("ta2", 3, 1, 1024), ("ta3", 3, 1, 65536), ("ta4", 16, 1, 65536),
          ("ta5", 3, 1024, 1024), ("ta6=cfg1", 3, 1024, 65536), ("ta7", 16, 1024, 65536),
          ("k8 m64 n2^18", 8, 64, 1 << 18), ("k16 m1024 n2^20", 16, 1024, 1 << 20), ("cfg2", 16, 4096, 1 << 20),
          ("k3 m8 n2^22", 3, 8, 1 << 22), ("cfg3", 8, 8, 1 << 26), ("cfg5/16", 3, 1 << 16, 1 << 20)]
rng = np.random.default_rng(5)
rows = []
nn.set_option("auto_gpus", 0)
for name, k, m, n in shapes:
    S = rng.random((m, k), dtype=np.float32)
    R = rng.random((n, k), dtype=np.float32)
    row = {"name": name, "k": k, "m": m, "n": n, "ms": {}}
    ref = None
    for g in counts:
        os.environ["NN_B200_GPUS"] = str(g)
        for _ in range(3):
            out = nn.cudaCallback(k, m, n, S, R)
        reps = 20 if m * n * k < 5e10 else 5
        ts = []
        for _ in range(reps):
            t0 = time.perf_counter()
            out = nn.cudaCallback(k, m, n, S, R)
            ts.append((time.perf_counter() - t0) * 1e3)
        ts.sort()
        row["ms"][str(g)] = round(ts[len(ts) // 2], 4)
        if ref is None:
            ref = out
        row.setdefault("same_result", True)
        row["same_result"] = row["same_result"] and bool(np.array_equal(out, ref))
    row["planned"] = nn.plan_gpus(k, m, n, vis)
    best = min(row["ms"], key=row["ms"].get)
    row["best"] = int(best)
    rows.append(row)
    print(f"{name:18s} " + " ".join(f"{g}:{row['ms'][str(g)]:9.3f}" for g in counts) +
          f"  planned {row['planned']} best {best} same {row['same_result']}", flush=True)
nn.set_option("auto_gpus", 1)
os.environ.pop("NN_B200_GPUS", None)
# with the planner in charge: never slower than one GPU
for row in rows:
    name, k, m, n = row["name"], row["k"], row["m"], row["n"]
    S = rng.random((m, k), dtype=np.float32)
    R = rng.random((n, k), dtype=np.float32)
    for _ in range(3):
        nn.cudaCallback(k, m, n, S, R)
    ts = []
    for _ in range(10 if m * n * k < 5e10 else 3):
        t0 = time.perf_counter()
        nn.cudaCallback(k, m, n, S, R)
        ts.append((time.perf_counter() - t0) * 1e3)
    ts.sort()
    row["ms_auto"] = round(ts[len(ts) // 2], 4)
    row["auto_used"] = nn.last_gpus()
json.dump({"visible": vis, "source": "scripts/gpu_count_sweep.py: wall clock of cudaCallback, malloc'ed inputs, median",
           "rows": rows}, open(os.path.join(ROOT, "gpurun_out", f"r02_gpu_count_{vis}.json"), "w"), indent=0)
print("auto:", [(r["name"], r["auto_used"], r["ms_auto"]) for r in rows])
