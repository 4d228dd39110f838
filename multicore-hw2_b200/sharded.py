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

import ctypes

from ._lib import check, lib
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


class PeerMerge:
    """The exchange step fused into the search kernels, one process per GPU (include/nn_b200.h, 2c):
    every rank's kernel folds its shard's candidates into rank 0's key array over NVLink peer memory
    (CUDA IPC + system-scope atomicMin) and rank 0's last CTA stores the merged indices -- one launch
    per rank and search instead of search -> all-reduce -> unpack.  torch.distributed is only used
    once, to exchange the 64-byte IPC handles.  Every rank must call :meth:`search` in step."""

    HANDLE_BYTES = 64

    def __init__(self, m: int, group=None, exchange: Optional[Callable] = None):
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.m = int(m)
        self._h = ctypes.c_void_p()
        check(lib().nn_b200_peer_create(self.m, self.rank, self.world, ctypes.byref(self._h)))
        mine = ctypes.create_string_buffer(self.HANDLE_BYTES)
        check(lib().nn_b200_peer_handle(self._h, mine, self.HANDLE_BYTES))
        if exchange is None:
            def exchange(b: bytes):
                if self.world == 1:
                    return [b]
                out = [None] * self.world
                dist.all_gather_object(out, b, group=group)
                return out
        handles = b"".join(exchange(mine.raw))
        if len(handles) != self.world * self.HANDLE_BYTES:
            raise ValueError("handle exchange returned the wrong number of bytes")
        check(lib().nn_b200_peer_attach(self._h, handles, len(handles)))
        if dist.is_initialized() and self.world > 1:
            dist.barrier(group=group)   # nobody searches before everybody is attached

    def search(self, S: torch.Tensor, R_shard: torch.Tensor, index_base: int, out: Optional[torch.Tensor] = None):
        """Fold this rank's shard; on rank 0 `out` (int32[m], device) holds the merged result once the
        kernel has completed.  Returns `out` (rank 0) or None."""
        k = S.shape[-1]
        m = S.numel() // k
        n = R_shard.numel() // k if R_shard.numel() else 0
        if self.rank == 0 and out is None:
            out = torch.empty(m, dtype=torch.int32, device=S.device)
        with torch.cuda.device(S.device):
            check(lib().nn_b200_peer_search(self._h, k, m, n, S.data_ptr(), R_shard.data_ptr() if n else None, index_base,
                                            out.data_ptr() if (self.rank == 0) else None,
                                            torch.cuda.current_stream().cuda_stream))
        return out if self.rank == 0 else None

    def error(self) -> bool:
        return bool(lib().nn_b200_peer_error(self._h))

    def close(self) -> None:
        if self._h:
            lib().nn_b200_peer_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
