"""Worker of tests/test_gpu_parity.py::test_peer_merge_between_processes (one process per GPU).
Every rank folds its shard of several different searches through multicore_hw2_b200.sharded.PeerMerge;
rank 0 compares every merged result with the CPU oracle.  Exit code 0 = all equal."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)


def main():
    import torch
    import torch.distributed as dist
    import cases
    import multicore_hw2_b200 as nn
    from multicore_hw2_b200 import sharded
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo")
    pm = sharded.PeerMerge(1200)
    shapes = [("duplicated", 16, 300, 40001), ("twins", 8, 200, 30011), ("quantized", 3, 1000, 5), ("uniform", 3, 1, 1),
              ("duplicated", 8, 8, 200003), ("specials", 5, 64, 9001), ("duplicated", 3, 1200, 70001),
              ("quantized", 16, 100, 50021), ("uniform", 7, 11, 65536)]
    bad = []
    outs = []
    data = []
    for i, (kind, k, m, n) in enumerate(shapes):
        S, R = cases.make(kind, 9000 + i, k, m, n)
        sh = sharded.ShardedSearch(n, rank, world)
        dS = torch.from_numpy(S).cuda()
        dR = torch.from_numpy(np.ascontiguousarray(sh.local_slice(R))).cuda() if sh.count else torch.empty((0, k), device="cuda")
        data.append((kind, k, m, n, S, R, dS, dR, sh.begin))
    dist.barrier()
    for rep in range(3):            # three rounds back to back: the two key buffers alternate and are re-used
        for kind, k, m, n, S, R, dS, dR, begin in data:
            out = pm.search(dS, dR, begin)
            outs.append((kind, k, m, n, S, R, out))
    torch.cuda.synchronize()
    if pm.error():
        bad.append("a device-side wait timed out")
    if rank == 0:
        from oracle import oracle
        for kind, k, m, n, S, R, out in outs:
            if not np.array_equal(out.cpu().numpy(), oracle.v0(S, R, threads=0)):
                bad.append((kind, k, m, n))
    dist.barrier()
    pm.close()
    dist.destroy_process_group()
    if bad:
        print("PEER MERGE MISMATCH", bad, flush=True)
        sys.exit(1)
    print(f"rank {rank}: peer merge ok", flush=True)


if __name__ == "__main__":
    main()
