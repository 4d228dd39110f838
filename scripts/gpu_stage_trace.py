"""Phase times of the staged pageable ingest (NN_B200_TRACE=1 prints them) for one 8 MiB reference set under
different staging set-ups: the open item 'a staged call carries ~1 ms of device time after everything is enqueued'."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import multicore_hw2_b200 as nn  # noqa: E402

k, m, n = 8, 64, 1 << 18
S = np.random.default_rng(1).random((m, k), dtype=np.float32)
R = np.random.default_rng(2).random((n, k), dtype=np.float32)


def med(reps=9):
    for _ in range(3):
        nn.search_host(S, R, k, num_gpus=1)
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        nn.search_host(S, R, k, num_gpus=1)
        ts.append((time.perf_counter() - t0) * 1e3)
    return sorted(ts)[len(ts) // 2]


nn.set_option("stage_min_bytes", 1 << 20)
for threads, chunk, one in [(8, 16 << 20, 1), (8, 16 << 20, 0), (1, 16 << 20, 1), (2, 16 << 20, 1), (1, 1 << 20, 1), (4, 1 << 20, 1), (8, 256 << 10, 1)]:
    nn.set_option("stage_threads", threads)
    nn.set_option("h2d_chunk_bytes", chunk)
    nn.set_option("stage_one_stream", one)
    print(f"threads={threads} chunk={chunk >> 10} KiB one_stream={one}: {med():.3f} ms", flush=True)
nn.set_option("stage_min_bytes", 1 << 40)
nn.set_option("stage_threads", -1)
nn.set_option("h2d_chunk_bytes", 16 << 20)
print(f"driver path: {med():.3f} ms", flush=True)
