"""World-size-2 (and 3) CPU test of the multi-GPU host logic over gloo: contiguous reference shards
(v8's partition, core.cu:875-883), global index bases, and the all-reduce(min) merge of packed keys
that replaces the reference's host-side second-level reduce (core.cu:936-957).  The per-shard search
is injected (the CPU oracle stands in for the CUDA kernel), so this runs without a GPU."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import cases


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _oracle_shard_search(S, R_shard, keys, index_base):
    """Stand-in for device.nearest_keys: fold the shard into int64 keys with global indices."""
    from oracle import oracle
    k = oracle.keys(S.numpy(), R_shard.numpy())
    k = (k & np.uint64(0xFFFFFFFF00000000)) | ((k & np.uint64(0xFFFFFFFF)) + np.uint64(index_base))
    # a shard in which nothing beat (INFINITY, 0) must not claim global index `index_base`
    untouched = (k >> np.uint64(32)) == np.uint64(0x7F800000)
    k[untouched] = np.uint64(0x7F80000000000000)
    torch.minimum(keys, torch.from_numpy(k.view(np.int64)), out=keys)
    return keys


def _worker(rank, world, port, kind, seed, k, m, n, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from multicore_hw2_b200 import sharded
        S, R = cases.make(kind, seed, k, m, n)
        sh = sharded.ShardedSearch(n, shard_search=_oracle_shard_search)
        assert (sh.rank, sh.world_size) == (rank, world)
        keys = torch.full((m,), 0x7F80000000000000, dtype=torch.int64)
        sh.keys(torch.from_numpy(S), torch.from_numpy(R[sh.begin:sh.begin + sh.count]), keys)
        if rank == 0:
            out.put((keys.numpy().view(np.uint64).copy(), sh.begin, sh.count))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,kind,k,m,n", [(2, "duplicated", 3, 64, 1001), (2, "twins", 16, 40, 777),
                                              (3, "quantized", 8, 33, 10), (2, "specials", 5, 16, 3)])
def test_sharded_merge_equals_single_search(oracle, world, kind, k, m, n):
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, kind, 4242, k, m, n, out)) for r in range(world)]
    for p in procs:
        p.start()
    keys, begin, count = out.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    S, R = cases.make(kind, 4242, k, m, n)
    want = oracle.keys(S, R)
    assert np.array_equal(keys, want)
    assert begin == 0 and count <= n


def test_merge_keys_is_noop_without_process_group():
    from multicore_hw2_b200 import sharded
    t = torch.tensor([5, 3], dtype=torch.int64)
    assert sharded.merge_keys(t) is t and t.tolist() == [5, 3]
