#!/bin/bash
# round 2: what the driver runs at round end on one GPU (tests, smoke, both bench arms), plus the pageable sizes
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
( time timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 ) 2>&1 | tail -9
( time python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 ) 2>&1 | tail -5
( time python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err ) 2>&1 | grep real
tail -c 300 gpurun_out/r02_bench_n1.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_bench_n1.json").read().strip().splitlines()[-1])
print("cfg4 value %.4e ms/step %.3f frac %.3f traffic %s e2e %.1f pinned %.1f index %.1f parity %s launches %d clocks %s" % (
    d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["traffic"], d["e2e"]["ms_per_call"], d["e2e"]["pinned"]["ms_per_call"],
    d["e2e"]["resident_index"]["ms_per_call"], d["parity_spot_check"], d["gpu_launches"], d["clocks"]))
print("cpu", {k: v for k, v in d["cpu_baseline"].items() if k != "sample"})
for k, v in (d["all_configs"] or {}).items():
    if "ms_per_step" in v:
        print(k, "ms %.4f frac %.3f e2e %.3f ms parity %s" % (v["ms_per_step"], v["roofline"]["frac"], v["e2e"]["ms_per_call"], v["parity_spot_check"]))
    else:
        print(k, json.dumps(v)[:600])
PY
python - <<'PY'
import sys, time, numpy as np, torch
sys.path.insert(0, ".")
import multicore_hw2_b200 as nn
for k, m, n in [(16, 1024, 1 << 16), (8, 64, 1 << 18), (16, 1024, 1 << 18), (16, 1024, 1 << 19), (16, 4096, 1 << 20), (8, 512, 1 << 23)]:
    S = np.random.default_rng(1).random((m, k), dtype=np.float32)
    R = np.random.default_rng(2).random((n, k), dtype=np.float32)
    Sp, Rp = torch.from_numpy(S).pin_memory(), torch.from_numpy(R).pin_memory()
    res = {}
    for name, fn in [("pageable", lambda: nn.search_host(S, R, k, num_gpus=1)), ("pinned", lambda: nn.search_host(Sp, Rp, k, num_gpus=1))]:
        for _ in range(3): fn()
        ts = []
        for _ in range(15):
            t0 = time.perf_counter(); fn(); ts.append((time.perf_counter() - t0) * 1e3)
        res[name] = sorted(ts)[len(ts) // 2]
    nn.set_option("stage_threads", 0)
    for _ in range(3): nn.search_host(S, R, k, num_gpus=1)
    ts = []
    for _ in range(15):
        t0 = time.perf_counter(); nn.search_host(S, R, k, num_gpus=1); ts.append((time.perf_counter() - t0) * 1e3)
    nn.set_option("stage_threads", -1)
    print(f"k={k} m={m} n={n} ({n*k*4/2**20:.0f} MiB): pageable {res['pageable']:.3f} ms (driver path {sorted(ts)[7]:.3f}) pinned {res['pinned']:.3f} ms ratio {res['pageable']/res['pinned']:.3f}")
PY
( time python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_bench_ref_n1.json 2> gpurun_out/r02_bench_ref_n1.err ) 2>&1 | grep real
cut -c1-400 gpurun_out/r02_bench_ref_n1.json
