import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import multicore_hw2_b200 as nn
import torch
def med(fn, reps=11):
    for _ in range(3): fn()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); fn(); ts.append((time.perf_counter() - t0) * 1e3)
    return sorted(ts)[len(ts) // 2]
for k, m, n in [(8, 64, 1 << 18), (16, 1024, 1 << 19), (16, 4096, 1 << 20), (8, 512, 1 << 23), (16, 65536, 1 << 24)]:
    S = np.random.default_rng(1).random((m, k), dtype=np.float32)
    R = np.random.default_rng(2).random((n, k), dtype=np.float32)
    Sp, Rp = torch.from_numpy(S).pin_memory(), torch.from_numpy(R).pin_memory()
    reps = 3 if m * n > 1e11 else 11
    out = {}
    nn.set_option("stage_min_bytes", 1 << 20)
    for one in (0, 1):
        nn.set_option("stage_one_stream", one)
        for grp in (1, 8):
            nn.set_option("search_group", grp)
            out[f"staged one_stream={one} group={grp}"] = med(lambda: nn.search_host(S, R, k, num_gpus=1), reps)
    nn.set_option("stage_one_stream", 0)
    nn.set_option("stage_min_bytes", 1 << 40)
    out["driver path"] = med(lambda: nn.search_host(S, R, k, num_gpus=1), reps)
    for grp in (1, 8):
        nn.set_option("search_group", grp)
        out[f"pinned group={grp}"] = med(lambda: nn.search_host(Sp, Rp, k, num_gpus=1), reps)
    print(f"k={k} m={m} n={n} ({n*k*4/2**20:.0f} MiB): " + "; ".join(f"{a} {b:.3f}" for a, b in out.items()), flush=True)
