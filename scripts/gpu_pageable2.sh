#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
nproc; lscpu | grep -E "Model name|Socket|NUMA node\(s\)|Thread" | head -5
python - <<'PY'
import time, numpy as np, torch, multicore_hw2_b200 as nn
rng = np.random.default_rng(1)
k, m, n = 8, 8, 1 << 26
S = rng.random((m, k), dtype=np.float32)
R = rng.random((n, k), dtype=np.float32)
out = np.empty(m, np.int32)
t0 = time.perf_counter(); R2 = R.copy(); print(f"numpy copy of 2 GiB on one thread: {R.nbytes/1e9/(time.perf_counter()-t0):.1f} GB/s"); del R2
for chunk in (4 << 20, 16 << 20, 64 << 20):
    for th in (0, 2, 4, 6, 8):
        nn.set_option("h2d_chunk_bytes", chunk); nn.set_option("stage_threads", th)
        for _ in range(2): nn.search_host(S, R, k, num_gpus=1, out=out)
        ts = []
        for _ in range(4):
            t0 = time.perf_counter(); nn.search_host(S, R, k, num_gpus=1, out=out); ts.append(time.perf_counter() - t0)
        print(f"chunk {chunk>>20} MiB threads {th}: median {1e3*np.median(ts):.1f} ms ({R.nbytes/1e9/np.median(ts):.1f} GB/s)")
PY
