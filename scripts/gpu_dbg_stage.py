import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import multicore_hw2_b200 as nn
k, m, n = 8, 64, 1 << 18
S = np.random.default_rng(1).random((m, k), dtype=np.float32)
R = np.random.default_rng(2).random((n, k), dtype=np.float32)
for i in range(4):
    t0 = time.perf_counter(); nn.search_host(S, R, k, num_gpus=1); print("call", i, (time.perf_counter() - t0) * 1e3, "ms", file=sys.stderr, flush=True)
