"""Staging-thread count for pageable reference sets of 64 MiB .. 1 GiB (ms per nn_b200_search_host call, one GPU)."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import multicore_hw2_b200 as nn  # noqa: E402


def med(fn, reps):
    for _ in range(2):
        fn()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        ts.append((time.perf_counter() - t0) * 1e3)
    return sorted(ts)[len(ts) // 2]


for k, m, n in [(16, 4096, 1 << 20), (8, 512, 1 << 22), (8, 512, 1 << 23), (8, 8, 1 << 26)]:
    S = np.random.default_rng(1).random((m, k), dtype=np.float32)
    R = np.random.default_rng(2).random((n, k), dtype=np.float32)
    out = {}
    for t in (1, 2, 4, 8):
        nn.set_option("stage_threads", t)
        out[f"{t} threads"] = med(lambda: nn.search_host(S, R, k, num_gpus=1), 7)
    nn.set_option("stage_threads", 0)
    out["driver"] = med(lambda: nn.search_host(S, R, k, num_gpus=1), 5)
    nn.set_option("stage_threads", -1)
    Sp, Rp = torch.from_numpy(S).pin_memory(), torch.from_numpy(R).pin_memory()
    out["pinned"] = med(lambda: nn.search_host(Sp, Rp, k, num_gpus=1), 7)
    print(f"k={k} m={m} n={n} ({n*k*4 >> 20} MiB): " + "; ".join(f"{a} {b:.2f}" for a, b in out.items()), flush=True)
