import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import cases
import multicore_hw2_b200 as nn
from oracle import oracle
order = sys.argv[1] if len(sys.argv) > 1 else "full"
seq = [("duplicated", 16, 300, 40001), ("twins", 8, 200, 30011), ("quantized", 3, 1000, 5), ("uniform", 3, 1, 1)]
if order == "only":
    seq = [("quantized", 3, 1000, 5)]
for rep in range(3):
    for kind, k, m, n in seq:
        S, R = cases.make(kind, 7000 + k, k, m, n)
        want = oracle.v0(S, R, threads=0)
        for p2p in (1, 0):
            nn.set_option("p2p_merge", p2p)
            got = nn.search_host(S, R, num_gpus=2)
            bad = np.nonzero(got != want)[0]
            print(rep, kind, "p2p", p2p, "mismatches", len(bad), "first", bad[:8], "got", got[bad[:8]], "want", want[bad[:8]], flush=True)
