#!/bin/bash
# Host entry with pageable (malloc) vs pinned host buffers.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "pageable or edge or ta_samples" 2>&1 | tail -3
python - <<'PY'
import time, numpy as np, torch, multicore_hw2_b200 as nn
rng = np.random.default_rng(1)
for (k, m, n) in [(16, 4096, 1 << 20), (8, 8, 1 << 26), (16, 8, 1 << 25)]:
    S = rng.random((m, k), dtype=np.float32)
    R = rng.random((n, k), dtype=np.float32)
    Sp, Rp = torch.from_numpy(S).pin_memory(), torch.from_numpy(R).pin_memory()
    out = np.empty(m, np.int32)
    for name, (a, b) in {"pageable": (S, R), "pinned": (Sp, Rp)}.items():
        for _ in range(2): nn.search_host(a, b, k, num_gpus=1, out=out)
        ts = []
        for _ in range(5):
            t0 = time.perf_counter(); nn.search_host(a, b, k, num_gpus=1, out=out); ts.append(time.perf_counter() - t0)
        gb = R.nbytes / 1e9
        print(f"k={k} m={m} n={n} {name}: median {1e3*np.median(ts):.3f} ms  ({gb/np.median(ts):.1f} GB/s of reference bytes)")
PY
