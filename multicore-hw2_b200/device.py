"""Device-resident building blocks on torch CUDA tensors.

torch is plumbing only: it owns the device memory and the stream; every kernel that runs is one of
this repository's sm_100a kernels, reached through the C ABI (include/nn_b200.h)."""
from __future__ import annotations

import torch

from ._lib import KEY_INIT, check, lib


def _stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def _chk(t: torch.Tensor, dtype, what: str) -> None:
    if not (t.is_cuda and t.is_contiguous() and t.dtype == dtype):
        raise ValueError(f"{what} must be a contiguous CUDA tensor of dtype {dtype}")


def new_keys(m: int, device=None) -> torch.Tensor:
    """Packed keys (d2 bits << 32 | index) in v0's start state (INFINITY, 0) (core.cu:39-40).
    Stored as int64: every key is < 2^63, so signed and unsigned order agree."""
    keys = torch.empty(m, dtype=torch.int64, device=device or "cuda")
    keys_init(keys)
    return keys


def keys_init(keys: torch.Tensor) -> None:
    _chk(keys, torch.int64, "keys")
    with torch.cuda.device(keys.device):
        check(lib().nn_b200_keys_init(keys.data_ptr(), keys.numel(), _stream_ptr()))


def nearest_keys(S: torch.Tensor, R: torch.Tensor, keys: torch.Tensor, index_base: int = 0) -> torch.Tensor:
    """Fold the nearest reference of every query of S [m][k] among R [n][k] into `keys`
    (fused distance + argmin + 64-bit atomicMin; replaces cudaCallbackKernel, core.cu:808-855)."""
    _chk(S, torch.float32, "S")
    _chk(R, torch.float32, "R")
    _chk(keys, torch.int64, "keys")
    k = S.shape[-1]
    m = S.numel() // k
    n = R.numel() // k if R.numel() else 0
    if R.numel() and R.shape[-1] != k:
        raise ValueError("S and R disagree on k")
    if keys.numel() != m:
        raise ValueError("keys must have one entry per query")
    with torch.cuda.device(S.device):
        check(lib().nn_b200_nearest_keys(k, m, n, S.data_ptr(), R.data_ptr(), index_base, keys.data_ptr(),
                                         _stream_ptr()))
    return keys


def nearest_keys_soa(S: torch.Tensor, R_soa: torch.Tensor, keys: torch.Tensor, index_base: int = 0) -> torch.Tensor:
    """Same search over references repacked to SoA [k][n] (the layout v4/v7/v8 search in)."""
    _chk(S, torch.float32, "S")
    _chk(R_soa, torch.float32, "R_soa")
    _chk(keys, torch.int64, "keys")
    k = S.shape[-1]
    m = S.numel() // k
    n = R_soa.numel() // k
    with torch.cuda.device(S.device):
        check(lib().nn_b200_nearest_keys_soa(k, m, n, S.data_ptr(), R_soa.data_ptr(), index_base, keys.data_ptr(),
                                             _stream_ptr()))
    return keys


def keys_unpack(keys: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
    _chk(keys, torch.int64, "keys")
    if out is None:
        out = torch.empty(keys.numel(), dtype=torch.int32, device=keys.device)
    _chk(out, torch.int32, "out")
    with torch.cuda.device(keys.device):
        check(lib().nn_b200_keys_unpack(keys.data_ptr(), keys.numel(), out.data_ptr(), _stream_ptr()))
    return out


def repack_soa(R: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
    """AoS [n][k] -> SoA [k][n] (replaces mat_inv_kernel, core.cu:792-807)."""
    _chk(R, torch.float32, "R")
    n, k = R.shape
    if out is None:
        out = torch.empty((k, n), dtype=torch.float32, device=R.device)
    _chk(out, torch.float32, "out")
    with torch.cuda.device(R.device):
        check(lib().nn_b200_repack_soa(k, n, R.data_ptr(), out.data_ptr(), _stream_ptr()))
    return out


class Workspace:
    """State of the one-launch search (include/nn_b200.h, section 2a): ticket counters + packed keys
    for up to `m` queries, initialised once; every search leaves it initialised."""

    def __init__(self, m: int, device=None):
        self.m = int(m)
        nbytes = lib().nn_b200_workspace_bytes(self.m)
        self.buf = torch.empty((nbytes + 7) // 8, dtype=torch.int64, device=device or "cuda")
        with torch.cuda.device(self.buf.device):
            check(lib().nn_b200_workspace_init(self.buf.data_ptr(), self.m, _stream_ptr()))

    @property
    def keys(self) -> torch.Tensor:
        """The workspace's key array as an int64 view (fold pieces into it with nearest_keys)."""
        off = lib().nn_b200_workspace_keys(self.buf.data_ptr()) - self.buf.data_ptr()
        return self.buf[off // 8: off // 8 + self.m]

    def finish(self, out: torch.Tensor | None = None, keys_out: torch.Tensor | None = None, m: int | None = None):
        m = self.m if m is None else m
        if out is None and keys_out is None:
            out = torch.empty(m, dtype=torch.int32, device=self.buf.device)
        with torch.cuda.device(self.buf.device):
            check(lib().nn_b200_workspace_finish(self.buf.data_ptr(), m, out.data_ptr() if out is not None else None,
                                                 keys_out.data_ptr() if keys_out is not None else None, _stream_ptr()))
        return out if out is not None else keys_out


def search(S: torch.Tensor, R: torch.Tensor, ws: Workspace | None = None, out: torch.Tensor | None = None,
           keys_out: torch.Tensor | None = None, index_base: int = 0) -> torch.Tensor:
    """Device-resident equivalent of one cudaCallback in ONE kernel launch: int32[m] nearest indices
    (and/or the final packed keys in `keys_out`).  Replaces transpose + cudaCallbackKernel + host reduce
    (core.cu:726-787)."""
    _chk(S, torch.float32, "S")
    _chk(R, torch.float32, "R")
    k = S.shape[-1]
    m = S.numel() // k
    n = R.numel() // k if R.numel() else 0
    if R.numel() and R.shape[-1] != k:
        raise ValueError("S and R disagree on k")
    if ws is None:
        ws = Workspace(m, S.device)
    if ws.m < m:
        raise ValueError("workspace too small for this many queries")
    if out is None and keys_out is None:
        out = torch.empty(m, dtype=torch.int32, device=S.device)
    if out is not None:
        _chk(out, torch.int32, "out")
    if keys_out is not None:
        _chk(keys_out, torch.int64, "keys_out")
    with torch.cuda.device(S.device):
        check(lib().nn_b200_search_device(k, m, n, S.data_ptr(), R.data_ptr(), index_base, ws.buf.data_ptr(),
                                          out.data_ptr() if out is not None else None,
                                          keys_out.data_ptr() if keys_out is not None else None, _stream_ptr()))
    return out if out is not None else keys_out


__all__ = ["KEY_INIT", "new_keys", "keys_init", "nearest_keys", "nearest_keys_soa", "keys_unpack", "repack_soa",
           "search", "Workspace"]
