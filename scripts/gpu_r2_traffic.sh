#!/bin/bash
# DRAM traffic of the config-4 kernel after the CTA re-ordering (one metric pass, full size) + timing check + pageable sizes
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
B=./multicore-hw2_b200/nn_bench
for a in "--k 16 --m 65536 --n 16777216" "--k 16 --m 4096 --n 1048576" "--k 3 --m 1048576 --n 1048576" "--k 3 --m 1024 --n 65536" "--k 8 --m 100 --n 4194304"; do
  $B $a --fused 1 --iters 3 --warmup 1 --check 0 | grep -v '"device"' | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(f\"k={d['k']:2d} m={d['m']:7d} n={d['n']:9d} {d['ms_med']*1e3:10.1f} us fp32 {d['fp32_frac_maxclk']:.3f} | {d['plan'][:70]}\")"
done
C="$B --k 16 --m 65536 --n 16777216 --fused 1 --iters 1 --warmup 1"
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:nn_qreg -s 1 -c 1 --csv --log-file gpurun_out/r02_cfg4_traffic.csv $C > /dev/null 2>&1
cat gpurun_out/r02_cfg4_traffic.csv | tail -5 | cut -c1-400
python - <<'PY'
import sys, time, numpy as np, torch
sys.path.insert(0, ".")
import multicore_hw2_b200 as nn
for k, m, n in [(16, 1024, 1 << 16), (8, 64, 1 << 18), (16, 1024, 1 << 18), (16, 4096, 1 << 20), (8, 512, 1 << 23)]:
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
    print(f"k={k} m={m} n={n} ({n*k*4/2**20:.0f} MiB): pageable {res['pageable']:.3f} ms pinned {res['pinned']:.3f} ms ratio {res['pageable']/res['pinned']:.3f}")
PY
