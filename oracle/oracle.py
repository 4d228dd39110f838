"""oracle/oracle.py -- TEST INFRASTRUCTURE ONLY.

ctypes front-end for the CPU checker:

* ``liboracle.so``  -- our restatement of ``v0::cudaCallback``
  (/root/reference/sources/src/core.cu:27-62) and of the TA generator
  (/root/reference/sources/src/generator.h:14-50), built by ``oracle/Makefile`` with
  ``-O2 -ffp-contract=off``.
* ``_ref/libref_v0*.so`` -- the REFERENCE's own ``v0`` compiled from /root/reference
  (strict flags / the reference's ``-Ofast`` flags), when it has been built.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU legs may import this
module.  The product package ``multicore-hw2_b200`` never does.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_f32p = ctypes.POINTER(ctypes.c_float)
_i32p = ctypes.POINTER(ctypes.c_int)
_u64p = ctypes.POINTER(ctypes.c_uint64)


def build(ref: bool = True, ref_gpu: bool = False, harness: bool = False) -> None:
    """Compile liboracle.so and, when /root/reference is present, oracle/_ref."""
    targets = ["all"]
    if os.path.isdir("/root/reference/sources/src"):
        if ref:
            targets.append("ref")
        if ref_gpu:
            targets.append("ref_gpu")
        if harness:
            targets.append("harness")
    subprocess.run(["make", "-s", "-C", _HERE] + targets, check=True)


def _load(path: str) -> ctypes.CDLL:
    if not os.path.exists(path):
        raise FileNotFoundError(f"{path} missing: run `make -C oracle` (or __graft_entry__.build())")
    return ctypes.CDLL(path)


_lib = None


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            build(ref=False)
        L = _load(path)
        L.nn_oracle_v0.argtypes = [ctypes.c_int] * 3 + [_f32p, _f32p, _i32p, _f32p]
        L.nn_oracle_v0_mt.argtypes = [ctypes.c_int] * 3 + [_f32p, _f32p, _i32p, _f32p, ctypes.c_int]
        L.nn_oracle_v0_mt.restype = ctypes.c_int
        L.nn_oracle_keys.argtypes = [ctypes.c_int] * 3 + [_f32p, _f32p, _u64p]
        L.nn_oracle_repack_soa.argtypes = [ctypes.c_int, ctypes.c_int, _f32p, _f32p]
        L.nn_oracle_ta_sample.argtypes = [ctypes.c_int, ctypes.c_int, _f32p, _f32p]
        L.nn_oracle_ta_shape.argtypes = [ctypes.c_int, _i32p, _i32p, _i32p]
        L.nn_oracle_sqdist.argtypes = [ctypes.c_int, _f32p, _f32p]
        L.nn_oracle_sqdist.restype = ctypes.c_float
        _lib = L
    return _lib


def _f32(a, shape=None) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.float32)
    if shape is not None:
        a = a.reshape(shape)
    return a


def _ptr(a: np.ndarray, t):
    return a.ctypes.data_as(t)


def v0(S, R, k: int | None = None, threads: int = 1, want_dist: bool = False):
    """v0 restatement.  S is [m][k], R is [n][k] (AoS, float32).  Returns int32[m]
    (and float32[m] winning squared distances when ``want_dist``)."""
    S = _f32(S)
    R = _f32(R)
    if k is None:
        k = S.shape[-1]
    m = S.size // k
    n = R.size // k
    out = np.empty(m, dtype=np.int32)
    dist = np.empty(m, dtype=np.float32) if want_dist else None
    dp = _ptr(dist, _f32p) if want_dist else None
    if threads == 1:
        lib().nn_oracle_v0(k, m, n, _ptr(S, _f32p), _ptr(R, _f32p), _ptr(out, _i32p), dp)
    else:
        lib().nn_oracle_v0_mt(k, m, n, _ptr(S, _f32p), _ptr(R, _f32p), _ptr(out, _i32p), dp, threads)
    return (out, dist) if want_dist else out


def keys(S, R, k: int | None = None) -> np.ndarray:
    """Packed (d^2 bits << 32 | index) keys, all host threads."""
    S = _f32(S)
    R = _f32(R)
    if k is None:
        k = S.shape[-1]
    m = S.size // k
    n = R.size // k
    out = np.empty(m, dtype=np.uint64)
    lib().nn_oracle_keys(k, m, n, _ptr(S, _f32p), _ptr(R, _f32p), _ptr(out, _u64p))
    return out


def repack_soa(R, k: int | None = None) -> np.ndarray:
    """AoS [n][k] -> SoA [k][n] (mat_inv_kernel, core.cu:792-807)."""
    R = _f32(R)
    if k is None:
        k = R.shape[-1]
    n = R.size // k
    out = np.empty((k, n), dtype=np.float32)
    lib().nn_oracle_repack_soa(k, n, _ptr(R, _f32p), _ptr(out, _f32p))
    return out


TA_SEED = 1000  # main.cu:43


def ta_shape(sample: int):
    k, m, n = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    lib().nn_oracle_ta_shape(sample, ctypes.byref(k), ctypes.byref(m), ctypes.byref(n))
    return k.value, m.value, n.value


def ta_sample(sample: int, seed: int = TA_SEED):
    """(S[m][k], R[n][k]) of TA sample `sample` (main.cu:28-39) under srand(seed)."""
    k, m, n = ta_shape(sample)
    S = np.empty((m, k), dtype=np.float32)
    R = np.empty((n, k), dtype=np.float32)
    lib().nn_oracle_ta_sample(seed, sample, _ptr(S, _f32p), _ptr(R, _f32p))
    return S, R


# ---- the reference's own v0, compiled from /root/reference (oracle/_ref) -----------------

_ref = {}


def ref_available(fast: bool = False) -> bool:
    return os.path.exists(os.path.join(_HERE, "_ref", "libref_v0_fast.so" if fast else "libref_v0.so"))


def ref_lib(fast: bool = False) -> ctypes.CDLL:
    if fast not in _ref:
        L = _load(os.path.join(_HERE, "_ref", "libref_v0_fast.so" if fast else "libref_v0.so"))
        L.ref_v0.argtypes = [ctypes.c_int] * 3 + [_f32p, _f32p, _i32p]
        L.ref_v0_mt.argtypes = [ctypes.c_int] * 3 + [_f32p, _f32p, _i32p, ctypes.c_int]
        L.ref_v0_mt.restype = ctypes.c_int
        _ref[fast] = L
    return _ref[fast]


def ref_v0(S, R, k: int | None = None, threads: int = 1, fast: bool = False):
    """The reference's unmodified v0::cudaCallback (core.cu:27-62).  `fast` selects the build
    with the reference's own host flags (-Ofast).  Returns (int32[m], threads_used)."""
    S = _f32(S)
    R = _f32(R)
    if k is None:
        k = S.shape[-1]
    m = S.size // k
    n = R.size // k
    out = np.empty(m, dtype=np.int32)
    L = ref_lib(fast)
    if threads == 1:
        L.ref_v0(k, m, n, _ptr(S, _f32p), _ptr(R, _f32p), _ptr(out, _i32p))
        used = 1
    else:
        used = L.ref_v0_mt(k, m, n, _ptr(S, _f32p), _ptr(R, _f32p), _ptr(out, _i32p), threads)
    return out, used
