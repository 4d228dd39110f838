"""Multi-GPU path, one process per GPU (the analogue of v8::cudaCallback, core.cu:856-958).

The reference shards the reference set contiguously over the GPUs of one process with OpenMP
threads (core.cu:873-883), gathers per-GPU candidate indices on the host and re-evaluates them
there (core.cu:925-957).  Here every rank owns one shard, produces per-query packed keys
``(float_bits(d2) << 32) | global_index`` on its GPU and the ranks are merged by ONE
all-reduce(min) over NCCL/NVLink.  Keys are < 2^63, so the int64 minimum torch.distributed
computes is the uint64 minimum the C ABI defines (ncclMin on ncclUint64 in nn_host.cu).

The host-side logic (shard arithmetic, global index bases, merge) is backend-agnostic: the CPU
tests run it with world_size 2 over gloo with an injected shard-search function."""
from __future__ import annotations

from typing import Callable, Optional

import torch
import torch.distributed as dist

from .api import shard_range


def merge_keys(keys: torch.Tensor, group=None) -> torch.Tensor:
    """In-place all-reduce(min) of packed keys across the ranks; no-op without a process group."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(keys, op=dist.ReduceOp.MIN, group=group)
    return keys


class ShardedSearch:
    """One rank's view of a reference set sharded over the ranks of a process group.

    `shard_search(S, R_shard, keys, index_base)` folds the shard into the keys; by default it is
    the CUDA path (:func:`multicore_hw2_b200.device.nearest_keys`)."""

    def __init__(self, n_total: int, rank: Optional[int] = None, world_size: Optional[int] = None, group=None,
                 shard_search: Optional[Callable] = None):
        if rank is None:
            rank = dist.get_rank(group) if dist.is_initialized() else 0
        if world_size is None:
            world_size = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank, self.world_size, self.group = rank, world_size, group
        self.n_total = n_total
        self.begin, self.count = shard_range(n_total, world_size, rank)
        if shard_search is None:
            from .device import nearest_keys
            shard_search = nearest_keys
        self._search = shard_search

    def local_slice(self, R_full):
        """This rank's rows of a full [n][k] reference array (view, no copy)."""
        return R_full[self.begin:self.begin + self.count]

    def keys(self, S: torch.Tensor, R_shard: torch.Tensor, keys: torch.Tensor) -> torch.Tensor:
        """Search the local shard (global indices) and merge across ranks.  `keys` must be in the
        start state (device.new_keys)."""
        if R_shard.shape[0] != self.count:
            raise ValueError(f"rank {self.rank} expects {self.count} reference rows, got {R_shard.shape[0]}")
        if self.count > 0:
            self._search(S, R_shard, keys, self.begin)
        return merge_keys(keys, self.group)
