"""CPU-side checks: the C-ABI library loads and exports exactly what include/nn_b200.h declares,
host-side helpers behave, the product never touches oracle/, and without a GPU the compute entry
points fail loudly (no fallback)."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def nn():
    import multicore_hw2_b200 as nn
    if not os.path.exists(nn.LIB_PATH):
        nn.build()
    nn.lib()
    return nn


def test_library_exports_every_declared_symbol(nn):
    hdr = open(os.path.join(ROOT, "include", "nn_b200.h")).read()
    declared = set(re.findall(r"NN_B200_API\s+[\w\s\*]+?\b(nn_b200_\w+)\s*\(", hdr))
    from multicore_hw2_b200 import _lib
    assert declared == set(_lib.SIGNATURES), "python signatures and header disagree"
    L = ctypes.CDLL(nn.LIB_PATH)
    for name in declared:
        assert getattr(L, name) is not None
    # the reference's own C++-linkage entry point (core.h:71) must be there for main.cu to link
    assert getattr(L, _lib.CXX_SYMBOL) is not None
    out = subprocess.run(["nm", "-D", "--defined-only", nn.LIB_PATH], capture_output=True, text=True).stdout
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l}
    assert exported == declared | {_lib.CXX_SYMBOL}, "unexpected exported symbols"


def test_no_contracted_multiply_add_in_device_code(nn):
    """v0's arithmetic is non-fused.  The shipped SASS may hold NO scalar FFMA at all.  The only
    FFMA2 allowed are the exact squares fma(d, d, -0.0) of the query-pair kernels (nn_kernels.cuh,
    sqdist_pair): there, per dimension, one FADD2 subtracts, one FFMA2 squares and (except for the
    first dimension) one FADD2 accumulates, so FADD2 : FFMA2 must be exactly (2k-1) : k and no
    FMUL2 may remain.  Had ptxas contracted a multiply into an add, the ratio would be off."""
    sass = subprocess.run(["cuobjdump", "-sass", nn.LIB_PATH], capture_output=True, text=True).stdout
    funcs = re.split(r"\n\s*Function : ", sass)[1:]
    assert len(funcs) > 100
    seen_pair_kernels = 0
    all_ops = set()
    for f in funcs:
        name = f.split("\n", 1)[0].strip()
        ops = re.findall(r"^\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_]+)", f, flags=re.M)
        all_ops.update(ops)
        n = {o: ops.count(o) for o in ("FFMA", "FFMA2", "FADD2", "FMUL2")}
        assert n["FFMA"] == 0, name
        mq = re.search(r"nn_qreg_kernelILi(\d+)ELi\d+ELi\d+ELi2E", name)
        mr = re.search(r"nn_(?:rreg|rtma|qflex)_kernelILi(\d+)E", name)
        if mq or mr:
            k = int((mq or mr).group(1))
            seen_pair_kernels += 1
            assert n["FFMA2"] > 0 and n["FMUL2"] == 0, name
            assert n["FADD2"] * k == n["FFMA2"] * (2 * k - 1), (name, n)
        else:
            assert n["FFMA2"] == 0, name
    assert seen_pair_kernels >= 14 * 2
    assert "UBLKCP" in all_ops                     # TMA bulk copy feeds the reference tiles
    assert "sm_100a" in subprocess.run(["cuobjdump", "-lelf", nn.LIB_PATH], capture_output=True, text=True).stdout


def test_shard_range_partitions_like_v8(nn):
    for n in (0, 1, 5, 1000, 65536, (1 << 24) + 3):
        for g in (1, 2, 3, 4, 8):
            got = [nn.shard_range(n, g, r) for r in range(g)]
            assert got[0][0] == 0 and sum(c for _, c in got) == n
            for (b0, c0), (b1, _) in zip(got, got[1:]):
                assert b0 + c0 == b1
            assert all(b % 4 == 0 or c == 0 for b, c in got)
            per = -(-n // g)
            assert max(c for _, c in got) <= per + 3   # core.cu:875's ceil(n/G), rounded to 4 points
    with pytest.raises(nn.NNError):
        nn.shard_range(10, 0, 0)


def test_no_gpu_means_loud_failure_not_fallback(nn):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(nn.NNError) as e:
        nn.search_host(np.zeros((2, 3), np.float32), np.ones((4, 3), np.float32))
    assert e.value.code == -4
    assert nn.device_count() == 0


def test_bad_arguments(nn):
    with pytest.raises(nn.NNError):
        nn.set_option("no_such_option", 1)
    with pytest.raises(ValueError):
        nn.cudaCallback(3, 2, 2, np.zeros(5, np.float32), np.zeros(6, np.float32))


def test_product_never_references_the_oracle():
    bad = []
    for base in ("multicore-hw2_b200", "include"):
        for dp, _, fs in os.walk(os.path.join(ROOT, base)):
            for f in fs:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", "Makefile", ".txt")):
                    txt = open(os.path.join(dp, f), errors="replace").read()
                    if re.search(r"(import\s+oracle|from\s+oracle|oracle/|liboracle|libref_|nn_oracle)", txt):
                        bad.append(os.path.join(dp, f))
    assert not bad, bad


def test_launch_planner_fills_the_gpu_without_shredding_the_work(nn):
    """The split planner (pure arithmetic) at the five BASELINE shapes on a 148-SM part, with the
    tile shapes and occupancies of nn_kernels.cuh: every split is a whole number of 4-reference
    chunks, the splits cover the reference set exactly once, the grid fills the SMs evenly (a
    whole number of CTAs per SM, or whole waves within 3%), and the work is never shredded into
    more than 8 waves of CTAs (a tie-break bug once produced one CTA per reference tile)."""
    import ctypes
    L = nn.lib()
    sms = 148
    # name: (k, queries/thread, CTAs/SM, m, n)
    shapes = {"cfg1 k3 q2": (3, 2, 9, 1024, 65536), "cfg1 k3 q4": (3, 4, 8, 1024, 65536),
              "cfg2 k16": (16, 4, 4, 4096, 1 << 20), "cfg4 k16": (16, 4, 4, 65536, 1 << 24),
              "cfg5 k3": (3, 8, 4, 1 << 20, 1 << 20), "tiny": (3, 1, 9, 1, 5), "one chunk": (8, 8, 4, 300, 4)}
    for name, (k, q, occ, m, n) in shapes.items():
        sp, rps = ctypes.c_int64(), ctypes.c_int64()
        assert L.nn_b200_plan_splits(k, q, occ, sms, m, n, ctypes.byref(sp), ctypes.byref(rps)) == 0
        qtiles = -(-m // (128 * q))
        total = sp.value * qtiles
        assert sp.value >= 1 and rps.value % 4 == 0 and rps.value >= 4, name
        assert sp.value * rps.value >= n > (sp.value - 1) * rps.value, (name, sp.value, rps.value)  # exact cover
        assert total <= 8 * sms * occ + qtiles, (name, total)
        if n * qtiles >= 64 * sms * occ:                              # enough work for a full grid
            per_sm = total / sms
            waves = -(-total // (sms * occ))
            even = abs(per_sm - round(per_sm)) < 0.03 * per_sm or total / (waves * sms * occ) > 0.97
            assert total >= sms and even, (name, total)


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the reference's v0 on the host cores) needs no GPU: one JSON line
    with the contract's keys, same metric/unit as the CUDA arm, zero launches."""
    import json
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--workload", "cfg1",
                          "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["impl"] == "reference" and line["unit"] == "pairs/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["gpu_launches"] == 0


def test_kernel_family_selection_rule(nn):
    """nn_b200_plan_variant (pure arithmetic): reference-register kernel up to 4 queries, reference-stream
    kernel for 5..24, then the phased query-register kernel wherever its groups x phases layout beats
    the query-register kernel's best padded tile, the query-register kernel otherwise -- checked against
    the B200 sweep profiles/r02_fewquery_crossover.json (nn_bench --sweep cross): the pick may never
    cost more than 20% over the fastest family and must be within 5% of it in 85% of the shapes."""
    import json
    L = nn.lib()
    assert L.nn_b200_plan_variant(8, 1, 1 << 26) == 2 and L.nn_b200_plan_variant(16, 4, 100) == 2
    assert L.nn_b200_plan_variant(8, 8, 1 << 26) == 4            # BASELINE config 3
    assert L.nn_b200_plan_variant(16, 4096, 1 << 20) == 1 and L.nn_b200_plan_variant(3, 1024, 65536) == 1
    assert L.nn_b200_plan_variant(3, 1 << 20, 1 << 20) == 1 and L.nn_b200_plan_variant(16, 65536, 1 << 24) == 1
    for k in (3, 8, 16):
        assert L.nn_b200_plan_variant(k, 100, 1 << 22) == 5      # 25 groups x 5 phases instead of a padded tile
        assert L.nn_b200_plan_variant(k, 256, 1 << 22) == 1      # exactly one 256-query tile: nothing to gain
    assert L.nn_b200_plan_variant(2, 8, 8) == -1
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    path = os.path.join(root, "profiles", "r02_fewquery_crossover.json")
    rows = json.load(open(path))["rows"]
    names = {1: "qreg_us", 4: "rtma_us", 5: "qflex_us"}
    good = 0
    for r in rows:
        pick = L.nn_b200_plan_variant(r["k"], r["m"], r["n"])
        assert pick in names and names[pick] in r, r
        best = min(v for kx, v in r.items() if kx in names.values())
        assert r[names[pick]] <= 1.20 * best + 2.0, r            # (+2 us: event resolution on the 10 us shapes)
        good += r[names[pick]] <= 1.05 * best + 1.0
    assert len(rows) >= 100 and good >= 0.85 * len(rows), (good, len(rows))


def test_phased_layout_uses_the_lanes(nn):
    """nn_b200_plan_flex (pure arithmetic): groups x phases fits the 128 threads of a CTA, the tiles
    cover the queries, and from 25 queries on at least 80% of the lanes carry useful queries."""
    import ctypes
    L = nn.lib()
    for k in range(3, 17):
        for m in list(range(1, 140)) + [200, 257, 300, 500, 777, 1000, 5000]:
            v = [ctypes.c_int() for _ in range(5)]
            assert L.nn_b200_plan_flex(k, m, *[ctypes.byref(x) for x in v]) == 0
            q, ng, np_, T, tq = [x.value for x in v]
            assert q in (2, 4, 8) and 1 <= ng * np_ <= 128 and ng * q >= tq and T * tq >= m > (T - 1) * tq - T
            if m >= 25:
                assert m / (T * ng * q) * (ng * np_ / 128) >= 0.80, (k, m, q, ng, np_, T, tq)


def test_gpu_count_planner(nn):
    """nn_b200_plan_gpus (pure arithmetic): the analogue of the reference's small-n single-GPU shortcut
    (core.cu:865-872).  Tiny calls stay on one GPU, big ones take every GPU, never more GPUs than
    references, monotone in the work."""
    P = nn.plan_gpus
    for vis in (1, 2, 4, 8):
        assert P(3, 1, 2, vis) == 1 and P(3, 2, 8, vis) == 1 and P(3, 1024, 1024, vis) == 1   # TA samples 0, 1, 5
        assert P(3, 1024, 65536, vis) == 1                                                   # BASELINE config 1
        assert P(16, 65536, 1 << 24, vis) == vis and P(8, 8, 1 << 26, vis) == vis            # configs 4 and 3
        assert P(3, 1 << 20, 1 << 20, vis) == vis and P(16, 4096, 1 << 20, vis) == vis       # configs 5 and 2
        assert 1 <= P(16, 1024, 65536, vis) <= vis
    assert P(16, 4096, 3, 8) <= 3
    assert P(2, 1, 1, 8) == -1
    last = 1
    for n in (1 << 10, 1 << 14, 1 << 18, 1 << 22, 1 << 26):
        g = P(8, 64, n, 8)
        assert g >= last
        last = g


def test_search_launches_over_the_h2d_chunks(nn):
    """nn_b200_plan_search_groups (pure arithmetic): how the host entry groups the H2D chunks of a shard
    into search launches.  Every chunk is searched exactly once, in order; the first four chunks go one
    by one (early start); a copy-bound search keeps small fixed groups (at most 1/24 of the list, so the
    un-overlapped last launch stays short); a compute-bound one takes what must have landed, never a chunk
    beyond the pessimistic copy-ahead estimate."""
    G = nn.plan_search_groups
    ramp = [32768, 65536, 131072]                       # the pinned path's ramp up to 16 MiB chunks at k = 16
    cfg4 = ramp + [262144] * 63 + [16777216 - sum(ramp) - 63 * 262144]
    assert sum(cfg4) == 1 << 24
    ends = G(16, 65536, cfg4)
    assert ends[-1] == len(cfg4) and ends == sorted(set(ends)) and ends[:4] == [1, 2, 3, 4]
    assert len(ends) <= 8                                                   # was 34 launches with 2 chunks each
    ahead = 0.5 * (3.0 * 16 * 65536 / (0.9 * 37.2e12)) / (16 * 4 / 8e9)
    starts = [0] + ends[:-1]
    for b, e in zip(starts, ends):
        if b >= 4 and e - b > 2:
            assert sum(cfg4[:e]) <= sum(cfg4[:b]) * ahead                   # only chunks that must have landed
    # the same shard as one of eight GPUs' (the host's copy bandwidth is shared): fixed groups again
    shard = [16384, 32768] + [65536] * 31
    e8 = G(16, 65536, shard, gpus=8)
    assert e8 == list(range(1, len(shard) + 1)) and len(G(16, 65536, shard, gpus=1)) < len(e8)
    # copy-bound (few queries): fixed groups, many launches, each at most 1/24 of the list
    cfg3 = [65536] * 4 + [524288] * 127
    ends3 = G(8, 8, cfg3)
    assert ends3[-1] == len(cfg3) and ends3[:4] == [1, 2, 3, 4]
    assert max(e - b for b, e in zip([0] + ends3[:-1], ends3)) <= max(1, len(cfg3) // 24)
    # mid-size (config 2: 4096 queries do not outrun the copy): fixed groups too
    assert G(16, 4096, [65536] * 16) == list(range(1, 17))
    # one chunk, no chunk, bad arguments
    assert G(3, 1024, [65536]) == [1] and G(3, 1024, []) == []
    with pytest.raises(nn.NNError):
        G(2, 1, [4096])
    with pytest.raises(nn.NNError):
        G(3, 1, [4096, 0])
    with pytest.raises(nn.NNError):
        G(3, 1, [4096], gpus=0)


def test_peer_merge_needs_a_gpu_and_checks_its_arguments(nn):
    """nn_b200_peer_*: argument validation works everywhere; without a CUDA device creation fails loudly
    (there is no host-memory stand-in for the NVLink merge)."""
    import ctypes
    import torch
    L = nn.lib()
    h = ctypes.c_void_p()
    assert L.nn_b200_peer_create(10, 2, 2, ctypes.byref(h)) == -1        # rank outside the world
    assert L.nn_b200_peer_create(-1, 0, 1, ctypes.byref(h)) == -1
    assert L.nn_b200_peer_search(None, 3, 1, 1, None, None, 0, None, None) == -1
    if not torch.cuda.is_available():
        assert L.nn_b200_peer_create(10, 0, 1, ctypes.byref(h)) in (-2, -4)
        from multicore_hw2_b200 import sharded
        with pytest.raises(nn.NNError):
            sharded.PeerMerge(10)


def test_bench_helpers_on_the_cpu():
    """bench.py's host-side pieces that need no GPU: both arms name the workload identically (the driver
    compares the strings), the oracle spot check accepts the oracle's own answer and rejects a corrupted one,
    and the roofline arithmetic is the north_star definition (3k lane-ops per pair vs n*k*4 bytes)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    assert bench.workload_string("cfg4", "strong", 1) == bench.workload_string("cfg4", "strong", 8)
    assert "configs[3]" in bench.workload_string("cfg4", "strong", 8)
    from oracle import oracle
    rng = np.random.default_rng(3)
    S, R = rng.random((40, 5), dtype=np.float32), rng.random((3001, 5), dtype=np.float32)
    good = oracle.v0(S, R, threads=0)
    bad = good.copy()
    bad[-1] = (bad[-1] + 1) % 3001          # the last query is always among the checked rows
    ok, n = bench.spot_check(5, 40, S, R, {"good": good, "bad": bad})
    assert n == 16 and ok == {"good": True, "bad": False}
    peaks = {"hbm_gbs": 6550.0, "sm_max_mhz": 1965.0, "source": "test"}
    r = bench.roofline_of(16, 4096, 1 << 20, 5.9, peaks, 148, "qreg k=16", "cfgX")
    assert r["bound"] == "fp32" and abs(r["frac"] - (3 * 16 * 4096 * (1 << 20) / (148 * 128 * 1.965e9)) / 5.9e-3) < 1e-9
    r = bench.roofline_of(8, 1, 1 << 26, 0.33, peaks, 148, "rreg k=8", "cfgX")
    assert r["bound"] == "hbm" and abs(r["achieved"] - (1 << 26) * 8 * 4 / 0.33e-3 / 1e9) < 1e-6
