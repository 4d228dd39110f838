#!/usr/bin/env python
"""oracle/ref_gpu_time.py -- TEST / BASELINE INFRASTRUCTURE ONLY.

Times the REFERENCE's own GPU path on this box: oracle/_ref/libref_gpu.so is the reference's whole
core.cu compiled unmodified for sm_100a with its own flags (oracle/Makefile, target ref_gpu).  Its
global `cudaCallback` (core.cu:1282-1297) forwards to v8, which takes the single-GPU v7 branch for
n <= min(2^18, 1024 m) (core.cu:871-872) -- the only branch whose results the reference validates,
i.e. the TA samples.  Wall clock around the call with malloc'ed inputs, as main.cu:69-73 times it.

Run as a subprocess by bench.py (a baseline beside our number, never on the product path): the library
runs every version once at load time (static WarmUP, core.cu:1274) and aborts without a GPU.
Prints one JSON line:  {"samples": [{"sample": i, "k":, "m":, "n":, "ms_med":, "ms_best":, "mismatches_vs_v0":}, ...]}"""
import ctypes
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))


def main():
    from oracle import oracle
    path = os.path.join(HERE, "_ref", "libref_gpu.so")
    if not os.path.exists(path):
        print(json.dumps({"unavailable": "oracle/_ref/libref_gpu.so not built"}))
        return
    t0 = time.perf_counter()
    L = ctypes.CDLL(path)  # static WarmUP runs here
    load_s = time.perf_counter() - t0
    fn = getattr(L, "_Z12cudaCallbackiiiPfS_PPi")
    fp, ip = ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_int)
    fn.argtypes = [ctypes.c_int] * 3 + [fp, fp, ctypes.POINTER(ip)]
    fn.restype = None
    libc = ctypes.CDLL(None)
    libc.free.argtypes = [ctypes.c_void_p]
    out = []
    for sample in [int(a) for a in sys.argv[1:]] or [6, 7]:
        k, m, n = oracle.ta_shape(sample)
        S, R = oracle.ta_sample(sample)
        want = oracle.v0(S, R, threads=0)
        ts, got = [], None
        for it in range(12):
            res = ip()
            t = time.perf_counter()
            fn(k, m, n, S.ctypes.data_as(fp), R.ctypes.data_as(fp), ctypes.byref(res))
            ts.append((time.perf_counter() - t) * 1e3)
            got = np.ctypeslib.as_array(res, shape=(m,)).copy()
            libc.free(ctypes.cast(res, ctypes.c_void_p))
        ts = sorted(ts[2:])
        out.append({"sample": sample, "k": k, "m": m, "n": n, "ms_med": ts[len(ts) // 2], "ms_best": ts[0],
                    "mismatches_vs_v0": int((got != want).sum())})
    print(json.dumps({"load_s": load_s, "samples": out,
                      "what": "reference core.cu recompiled for sm_100a (-O3 -use_fast_math), global cudaCallback -> v8 -> v7, "
                              "wall clock per call, malloc'ed inputs"}))


if __name__ == "__main__":
    main()
