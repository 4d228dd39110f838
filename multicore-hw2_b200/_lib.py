"""Loader for the C-ABI shared library ``libnn_b200.so`` (include/nn_b200.h).

The library is built in-tree by ``multicore-hw2_b200/csrc/Makefile`` (``__graft_entry__.build()``).
There is no fallback of any kind: if the library is missing, or a call fails, an exception is
raised -- nothing in this package computes on the CPU."""
from __future__ import annotations

import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libnn_b200.so")

OK, EINVAL, ECUDA, ENCCL, ENODEV = 0, -1, -2, -3, -4
KEY_INIT = 0x7F80000000000000

_f32p = ctypes.POINTER(ctypes.c_float)
_i32p = ctypes.POINTER(ctypes.c_int)

# name -> (restype, argtypes); exactly the declarations of include/nn_b200.h
SIGNATURES = {
    "nn_b200_cudaCallback": (None, [ctypes.c_int] * 3 + [_f32p, _f32p, ctypes.POINTER(_i32p)]),
    "nn_b200_search_host": (ctypes.c_int, [ctypes.c_int] * 3 + [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]),
    "nn_b200_keys_init": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]),
    "nn_b200_nearest_keys": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p,
                                            ctypes.c_uint32, ctypes.c_void_p, ctypes.c_void_p]),
    "nn_b200_keys_unpack": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]),
    "nn_b200_repack_soa": (ctypes.c_int, [ctypes.c_int, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]),
    "nn_b200_nearest_keys_soa": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p,
                                                ctypes.c_uint32, ctypes.c_void_p, ctypes.c_void_p]),
    "nn_b200_workspace_bytes": (ctypes.c_size_t, [ctypes.c_int]),
    "nn_b200_workspace_init": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]),
    "nn_b200_workspace_keys": (ctypes.c_void_p, [ctypes.c_void_p]),
    "nn_b200_search_device": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p,
                                             ctypes.c_uint32, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                             ctypes.c_void_p]),
    "nn_b200_workspace_finish": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                                                ctypes.c_void_p]),
    "nn_b200_peer_create": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_void_p)]),
    "nn_b200_peer_handle": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t]),
    "nn_b200_peer_attach": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t]),
    "nn_b200_peer_search": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int64, ctypes.c_void_p,
                                           ctypes.c_void_p, ctypes.c_uint32, ctypes.c_void_p, ctypes.c_void_p]),
    "nn_b200_peer_error": (ctypes.c_int, [ctypes.c_void_p]),
    "nn_b200_peer_destroy": (None, [ctypes.c_void_p]),
    "nn_b200_shard_range": (ctypes.c_int, [ctypes.c_int64, ctypes.c_int, ctypes.c_int,
                                           ctypes.POINTER(ctypes.c_int64), ctypes.POINTER(ctypes.c_int64)]),
    "nn_b200_device_count": (ctypes.c_int, [ctypes.c_int64]),
    "nn_b200_launch_count": (ctypes.c_int64, []),
    "nn_b200_plan_gpus": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_int64, ctypes.c_int]),
    "nn_b200_plan_search_groups": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                                  ctypes.c_int, ctypes.c_void_p]),
    "nn_b200_last_gpus": (ctypes.c_int, []),
    "nn_b200_last_error": (ctypes.c_char_p, []),
    "nn_b200_set_option": (ctypes.c_int, [ctypes.c_char_p, ctypes.c_int64]),
    "nn_b200_probe_fp32": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_double)]),
    "nn_b200_plan_splits": (ctypes.c_int, [ctypes.c_int] * 4 + [ctypes.c_int64] * 2 + [ctypes.POINTER(ctypes.c_int64)] * 2),
    "nn_b200_index_create": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(ctypes.c_void_p)]),
    "nn_b200_index_search": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]),
    "nn_b200_index_info": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int64), ctypes.POINTER(ctypes.c_int)]),
    "nn_b200_index_destroy": (None, [ctypes.c_void_p]),
    "nn_b200_warmup": (ctypes.c_int, []),
    "nn_b200_plan_flex": (ctypes.c_int, [ctypes.c_int, ctypes.c_int] + [ctypes.POINTER(ctypes.c_int)] * 5),
    "nn_b200_plan_variant": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_int64]),
    "nn_b200_describe_plan": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_int64, ctypes.c_char_p, ctypes.c_size_t]),
}
CXX_SYMBOL = "_Z12cudaCallbackiiiPfS_PPi"  # ::cudaCallback(int,int,int,float*,float*,int**), core.h:71


class NNError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"nn_b200 error {code}: {msg}")
        self.code = code


def build(jobs: int = 8) -> None:
    """Compile libnn_b200.so (+ nn_bench) for sm_100a with nvcc."""
    subprocess.run(["make", "-s", f"-j{jobs}", "-C", os.path.join(_HERE, "csrc")], check=True)


_lib = None


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise FileNotFoundError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU or PyTorch fallback for this path)")
        L = ctypes.CDLL(LIB_PATH, mode=ctypes.RTLD_GLOBAL)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc: int) -> None:
    if rc != OK:
        raise NNError(rc, (lib().nn_b200_last_error() or b"").decode(errors="replace"))
