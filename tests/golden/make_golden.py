"""Regenerates tests/golden/*.  Run in the build container, where /root/reference exists:

    python tests/golden/make_golden.py

ta_results.json   <- /root/reference/results.csv: odd lines = nearest indices of TA samples 0..7
                     (seed 1000, /root/reference/sources/src/main.cu:28-43), even lines = distances.
ref_v0_cases.json <- outputs of the REFERENCE's own v0::cudaCallback
                     (/root/reference/sources/src/core.cu:27-62, compiled by oracle/Makefile into
                     oracle/_ref/libref_v0.so) on the seeded cases of tests/cases.py, with a CRC of
                     the inputs so RNG drift is caught.
The GPU box has no /root/reference; the committed fixtures are what travels."""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(HERE))

import cases  # noqa: E402
from oracle import oracle  # noqa: E402


def main():
    oracle.build(ref=True)
    lines = open("/root/reference/results.csv").read().strip().split("\n")
    assert len(lines) == 16
    idx = [[int(x) for x in l.strip().strip(",").split(",")] for l in lines[0::2]]
    dist = [[float(x) for x in l.strip().strip(",").split(",")] for l in lines[1::2]]
    json.dump({"source": "/root/reference/results.csv", "seed": 1000,
               "samples": [list(oracle.ta_shape(i)) for i in range(8)],
               "indices": idx, "distances": dist},
              open(os.path.join(HERE, "ta_results.json"), "w"))

    out = []
    for kind, seed, k, m, n in cases.GOLDEN_CASES:
        S, R = cases.make(kind, seed, k, m, n)
        res, _ = oracle.ref_v0(S, R, k)
        out.append({"kind": kind, "seed": seed, "k": k, "m": m, "n": n,
                    "crc": cases.checksum(S, R), "indices": res.tolist()})
    json.dump({"source": "oracle/_ref/libref_v0.so = /root/reference/sources/src/core.cu:25-63, "
                         "g++ -O2 -ffp-contract=off", "cases": out},
              open(os.path.join(HERE, "ref_v0_cases.json"), "w"))
    print("wrote", len(out), "reference cases and 8 TA samples")


if __name__ == "__main__":
    main()
