#!/bin/bash
# round 2: full GPU suite with the new tests, flex unroll A/B, bench line incl. reference GPU leg
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
( timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/r2_pytest.log
tail -4 gpurun_out/r2_pytest.log
for d in multicore-hw2_b200 build/alt_*; do
  [ -x $d/nn_bench ] || continue
  for a in "--k 3 --m 100" "--k 8 --m 100" "--k 16 --m 100" "--k 3 --m 32" "--k 8 --m 32" "--k 3 --m 200" "--k 8 --m 200"; do
    $d/nn_bench $a --n 4194304 --variant 5 --iters 9 2>/dev/null | grep -v '"device"' | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('$(basename $d)'.ljust(20), f\"k={d['k']:2d} m={d['m']:4d} {d['ms_med']*1e3:8.1f} us fp32 {d['fp32_frac_maxclk']:.3f} | {d['plan'][:60]}\")"
  done
done
( time python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_n1b.json 2> gpurun_out/r2_bench_n1b.err ) 2>&1 | grep real
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2_bench_n1b.json").read().strip().splitlines()[-1])
print("cfg4 ms/step %.3f frac %.3f e2e %.1f pinned %.1f" % (d["ms_per_step"], d["roofline"]["frac"], d["e2e"]["ms_per_call"], d["e2e"]["pinned"]["ms_per_call"]))
for k, v in (d["all_configs"] or {}).items():
    if "ms_per_step" in v:
        print(k, "ms %.4f frac %.3f e2e %.3f ms parity %s" % (v["ms_per_step"], v["roofline"]["frac"], v["e2e"]["ms_per_call"], v["parity_spot_check"]))
    else:
        print(k, v)
PY
python - <<'PY'
# pageable vs pinned through the host entry, cfg2 and a 16 MiB / 256 MiB set
import sys, time, numpy as np, torch
sys.path.insert(0, ".")
import multicore_hw2_b200 as nn
for k, m, n in [(16, 4096, 1 << 20), (16, 1024, 1 << 18), (8, 512, 1 << 23)]:
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
    print(f"k={k} m={m} n={n} ({n*k*4>>20} MiB): pageable {res['pageable']:.3f} ms pinned {res['pinned']:.3f} ms ratio {res['pageable']/res['pinned']:.3f}")
PY
