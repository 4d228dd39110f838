#!/bin/bash
# Multi-GPU host entry: parity with both merge modes, then call latency P2P-atomics vs NCCL.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "multi_gpu" > gpurun_out/pytest_multi.log 2>&1; tail -3 gpurun_out/pytest_multi.log
python - <<'PY'
import time, numpy as np, torch, multicore_hw2_b200 as nn
g = torch.cuda.device_count()
rng = np.random.default_rng(1)
for (k, m, n) in [(3, 1024, 65536), (8, 8, 1 << 24), (16, 4096, 1 << 20)]:
    S = torch.from_numpy(rng.random((m, k), dtype=np.float32)).pin_memory()
    R = torch.from_numpy(rng.random((n, k), dtype=np.float32)).pin_memory()
    ref = None
    for gpus in sorted({1, 2, g}):
        for p2p in (1, 0):
            if gpus == 1 and p2p == 0: continue
            nn.set_option("p2p_merge", p2p)
            out = np.empty(m, np.int32)
            for _ in range(3): nn.search_host(S, R, k, num_gpus=gpus, out=out)
            ts = []
            for _ in range(15):
                t0 = time.perf_counter(); nn.search_host(S, R, k, num_gpus=gpus, out=out); ts.append(time.perf_counter() - t0)
            if ref is None: ref = out.copy()
            print(f"k={k} m={m} n={n} gpus={gpus} merge={'p2p-atomics' if p2p else 'nccl'}: median {1e3*np.median(ts):.3f} ms  min {1e3*min(ts):.3f} ms  same={np.array_equal(out, ref)}")
PY
