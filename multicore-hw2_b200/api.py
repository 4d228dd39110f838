"""Host-pointer API: the reference's ``cudaCallback`` and its helpers, through the C ABI."""
from __future__ import annotations

import ctypes

import numpy as np

from . import _lib
from ._lib import check, lib

_libc = ctypes.CDLL(None)
_libc.free.argtypes = [ctypes.c_void_p]


def _host_f32(a, what: str) -> np.ndarray:
    a = np.asarray(a)
    if a.dtype != np.float32 or not a.flags["C_CONTIGUOUS"]:
        a = np.ascontiguousarray(a, dtype=np.float32)
    return a


def cudaCallback(k: int, m: int, n: int, searchPoints, referencePoints) -> np.ndarray:
    """``cudaCallback(k, m, n, searchPoints, referencePoints, &results)`` of the reference
    (core.h:71; core.cu:1282-1297), through the exported C symbol ``nn_b200_cudaCallback``.

    searchPoints is [m][k], referencePoints is [n][k] (host, float32, AoS).  Returns the int32[m]
    nearest indices.  The C side malloc()s the result exactly like the reference (core.cu:935);
    it is copied into a numpy array and freed here.  Failures terminate the process with the
    reference's "Error: ..." message (CHECK macro, core.h:77-87); use :func:`search_host` for an
    exception instead."""
    S = _host_f32(searchPoints, "searchPoints")
    R = _host_f32(referencePoints, "referencePoints")
    if S.size != k * m or R.size != k * n:
        raise ValueError(f"expected {k*m} query floats and {k*n} reference floats, got {S.size} and {R.size}")
    res = ctypes.POINTER(ctypes.c_int)()
    lib().nn_b200_cudaCallback(k, m, n, S.ctypes.data_as(ctypes.POINTER(ctypes.c_float)),
                               R.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), ctypes.byref(res))
    out = np.ctypeslib.as_array(res, shape=(max(m, 1),))[:m].copy() if m > 0 else np.empty(0, np.int32)
    _libc.free(ctypes.cast(res, ctypes.c_void_p))
    return out.astype(np.int32, copy=False)


def search_host(searchPoints, referencePoints, k: int | None = None, num_gpus: int = 0, out=None) -> np.ndarray:
    """Same search with an exception on failure.  Accepts numpy arrays or (pinned) torch CPU
    tensors; the host buffers are handed to the C ABI as they are (no copy when float32/contiguous)."""
    S, s_ptr = _as_host(searchPoints)
    R, r_ptr = _as_host(referencePoints)
    if k is None:
        k = int(S.shape[-1])
    m, n = _count(S, k), _count(R, k)
    if out is None:
        out = np.empty(m, dtype=np.int32)
    o, o_ptr = _as_host(out, dtype="int32", output=True)
    if _count(o, 1) < m:
        raise ValueError("`out` is smaller than the number of queries")
    check(lib().nn_b200_search_host(k, m, n, s_ptr, r_ptr, o_ptr, num_gpus))
    return out


def _as_host(a, dtype: str = "float32", output: bool = False):
    """(object kept alive, raw pointer) for a numpy array or a torch CPU tensor of exactly `dtype`.
    Inputs of another dtype or layout are converted (a copy); an OUTPUT buffer must already be a
    contiguous array of the right dtype -- results written into a silent copy would be lost."""
    if hasattr(a, "data_ptr"):  # torch tensor
        import torch
        if a.is_cuda:
            raise ValueError("host entry point takes host buffers; use multicore_hw2_b200.device for CUDA tensors")
        want = getattr(torch, dtype)
        if output:
            if a.dtype != want or not a.is_contiguous():
                raise ValueError(f"`out` must be a contiguous {dtype} tensor")
        else:
            a = a.to(want).contiguous()
        return a, a.data_ptr()
    if output:
        if not isinstance(a, np.ndarray) or a.dtype != np.dtype(dtype) or not a.flags["C_CONTIGUOUS"] \
                or not a.flags["WRITEABLE"]:
            raise ValueError(f"`out` must be a writeable C-contiguous numpy array of dtype {dtype}")
        return a, a.ctypes.data
    a = np.ascontiguousarray(a, dtype=dtype)
    return a, a.ctypes.data


def _count(a, k: int) -> int:
    return (a.numel() if hasattr(a, "numel") else a.size) // k


def shard_range(n: int, num_shards: int, shard: int):
    """(begin, count) of a contiguous reference shard (v8's partition, core.cu:875-883)."""
    b, c = ctypes.c_int64(), ctypes.c_int64()
    check(lib().nn_b200_shard_range(n, num_shards, shard, ctypes.byref(b), ctypes.byref(c)))
    return b.value, c.value


def device_count(n: int = 1 << 30) -> int:
    return lib().nn_b200_device_count(n)


def launch_count() -> int:
    return lib().nn_b200_launch_count()


def plan_gpus(k: int, m: int, n: int, visible: int) -> int:
    """GPUs a host-entry call would use out of `visible` (pure arithmetic; core.cu:865-872's job)."""
    return lib().nn_b200_plan_gpus(k, m, n, visible)


def plan_search_groups(k: int, m: int, chunk_refs, gpus: int = 1) -> list:
    """Host entry, ingest pipeline (pure arithmetic): for H2D chunks of `chunk_refs` references each (one
    shard of a call that drives `gpus` GPUs), the index one past the last chunk of every search launch."""
    refs = np.ascontiguousarray(chunk_refs, dtype=np.int64)
    ends = np.zeros(max(len(refs), 1), dtype=np.int32)
    n = lib().nn_b200_plan_search_groups(k, m, gpus, refs.ctypes.data, len(refs), ends.ctypes.data)
    if n < 0:
        check(n)
    return ends[:n].tolist()


def last_gpus() -> int:
    return lib().nn_b200_last_gpus()


def set_option(name: str, value: int) -> None:
    check(lib().nn_b200_set_option(name.encode(), int(value)))


def describe_plan(k: int, m: int, n: int) -> str:
    buf = ctypes.create_string_buffer(512)
    check(lib().nn_b200_describe_plan(k, m, n, buf, 512))
    return buf.value.decode()


def probe_fp32(mode: int = 0, iters: int = 20000) -> float:
    """Measured non-fused FP32 issue rate of the current device in lane-ops/s (roofline denominator).
    mode 0 scalar FMUL/FADD, 1 packed f32x2, 2..5 mixes (see include/nn_b200.h)."""
    v = ctypes.c_double()
    check(lib().nn_b200_probe_fp32(int(mode), iters, ctypes.byref(v)))
    return v.value


class Index:
    """Resident reference index (build once, query many): the reference set stays sharded in HBM and
    a search moves only the queries and the indices over PCIe (include/nn_b200.h, section 2b).
    Same results as :func:`search_host` on the same data."""

    def __init__(self, referencePoints, k: int | None = None, num_gpus: int = 0):
        R, r_ptr = _as_host(referencePoints)
        if k is None:
            k = int(R.shape[-1])
        n = _count(R, k)
        self._h = ctypes.c_void_p()
        check(lib().nn_b200_index_create(k, n, r_ptr, num_gpus, ctypes.byref(self._h)))
        self.k, self.n = k, n
        g = ctypes.c_int()
        check(lib().nn_b200_index_info(self._h, None, None, ctypes.byref(g)))
        self.gpus = g.value

    def search(self, searchPoints, out=None) -> np.ndarray:
        if not self._h:
            raise ValueError("index is closed")
        S, s_ptr = _as_host(searchPoints)
        m = _count(S, self.k)
        if out is None:
            out = np.empty(m, dtype=np.int32)
        o, o_ptr = _as_host(out, dtype="int32", output=True)
        if _count(o, 1) < m:
            raise ValueError("`out` is smaller than the number of queries")
        check(lib().nn_b200_index_search(self._h, m, s_ptr, o_ptr))
        return out

    def close(self) -> None:
        if self._h:
            lib().nn_b200_index_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
