"""Parity of the CUDA path (through the C ABI, libnn_b200.so) against the CPU oracle and the
committed reference fixtures.  Bit-exact: indices AND packed keys (distance bits) must be equal.
Run on the B200 box: python -m pytest tests -m gpu."""
import json
import os

import numpy as np
import pytest

import cases

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TA = json.load(open(os.path.join(GOLD, "ta_results.json")))
REF = json.load(open(os.path.join(GOLD, "ref_v0_cases.json")))

VARIANTS = {"auto": 0, "qreg": 1, "rreg": 2, "plain": 3, "rtma": 4, "qflex": 5}


@pytest.fixture(scope="module")
def nn():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("gpu test collected without a CUDA device")
    import multicore_hw2_b200 as nn
    nn.lib()
    yield nn
    nn.set_option("variant", 0)
    nn.set_option("qreg_q", 0)
    nn.set_option("math", 2)


def gpu_keys(nn, S, R, variant="auto", soa=False, index_base=0, q=0, math=2):
    import torch
    from multicore_hw2_b200 import device
    nn.set_option("variant", VARIANTS[variant])
    nn.set_option("qreg_q", q)
    nn.set_option("math", math)
    try:
        dS = torch.from_numpy(np.ascontiguousarray(S)).cuda()
        dR = torch.from_numpy(np.ascontiguousarray(R)).cuda()
        keys = device.new_keys(S.shape[0])
        if soa:
            device.nearest_keys_soa(dS, device.repack_soa(dR), keys, index_base)
        else:
            device.nearest_keys(dS, dR, keys, index_base)
        torch.cuda.synchronize()
        return keys.cpu().numpy().view(np.uint64)
    finally:
        nn.set_option("variant", 0)
        nn.set_option("qreg_q", 0)
        nn.set_option("math", 2)


@pytest.mark.parametrize("variant", ["auto", "qreg", "rreg", "plain", "rtma", "qflex"])
@pytest.mark.parametrize("case", REF["cases"], ids=lambda c: f"{c['kind']}-k{c['k']}-m{c['m']}-n{c['n']}")
def test_reference_v0_fixture(nn, oracle, case, variant):
    """Indices equal the reference's own v0 outputs (fixture); keys equal the oracle's."""
    S, R = cases.make(case["kind"], case["seed"], case["k"], case["m"], case["n"])
    keys = gpu_keys(nn, S, R, variant)
    assert (keys & 0xFFFFFFFF).astype(np.int64).tolist() == case["indices"]
    assert np.array_equal(keys, oracle.keys(S, R))


@pytest.mark.parametrize("variant", ["auto", "qreg", "rreg", "plain", "rtma", "qflex"])
def test_one_launch_search_reuses_its_workspace(nn, oracle, variant):
    """nn_b200_search_device: search + merge + index store in one launch (the last CTA of every ticket
    group finishes and restores the workspace).  One workspace serves every fixture case in turn --
    any state left behind by a search would corrupt the next -- and returns indices, keys, or both."""
    import torch
    from multicore_hw2_b200 import device
    big = max(c["m"] for c in REF["cases"])
    ws = device.Workspace(big)
    nn.set_option("variant", VARIANTS[variant])
    try:
        for i, case in enumerate(REF["cases"]):
            S, R = cases.make(case["kind"], case["seed"], case["k"], case["m"], case["n"])
            dS, dR = torch.from_numpy(S).cuda(), torch.from_numpy(R).cuda()
            m = case["m"]
            out = torch.full((m,), -7, dtype=torch.int32, device="cuda")
            keys = torch.zeros(m, dtype=torch.int64, device="cuda")
            mode = i % 3
            device.search(dS, dR, ws, out=out if mode != 1 else None, keys_out=keys if mode != 0 else None)
            if mode != 1:
                assert out.cpu().tolist() == case["indices"], (case, variant)
            if mode != 0:
                assert np.array_equal(keys.cpu().numpy().view(np.uint64), oracle.keys(S, R)), (case, variant)
        assert bool((ws.keys == device.KEY_INIT).all()) and int(ws.buf[:512].abs().sum()) == 0  # start state again
        # pieces folded into the workspace's keys, then one finishing kernel
        S, R = cases.make("duplicated", 77, 5, 300, 9001)
        dS, dR = torch.from_numpy(S).cuda(), torch.from_numpy(R).cuda()
        for b in (6000, 3000, 0):
            device.nearest_keys(dS, dR[b:b + 3001], ws.keys[:300], b)
        assert np.array_equal(ws.finish(m=300).cpu().numpy(), oracle.v0(S, R, threads=0))
        assert bool((ws.keys == device.KEY_INIT).all())
        # n = 0: v0's start state, index 0
        assert device.search(dS, dR[:0], ws).cpu().tolist() == [0] * 300
    finally:
        nn.set_option("variant", 0)


@pytest.mark.parametrize("case", [c for c in REF["cases"] if c["kind"] in ("twins", "quantized", "specials")],
                         ids=lambda c: f"{c['kind']}-k{c['k']}")
def test_soa_path_and_other_math_modes(nn, oracle, case):
    S, R = cases.make(case["kind"], case["seed"], case["k"], case["m"], case["n"])
    want = oracle.keys(S, R)
    assert np.array_equal(gpu_keys(nn, S, R, soa=True), want)
    for math in (0, 1):   # scalar / dimension-pair math: only in -DNN_AB_MATH builds of the library
        try:
            nn.set_option("math", math)
        except nn.NNError:
            continue
        assert np.array_equal(gpu_keys(nn, S, R, "qreg", math=math), want)
    for q in (1, 2):
        assert np.array_equal(gpu_keys(nn, S, R, "qreg", q=q), want)


@pytest.mark.parametrize("sample", range(8))
def test_ta_samples_through_cudaCallback(nn, oracle, sample):
    """The reference's eight TA samples (main.cu:28-39, seed 1000) through the drop-in entry
    point, against the reference's golden results.csv."""
    k, m, n = oracle.ta_shape(sample)
    S, R = oracle.ta_sample(sample)
    got = nn.cudaCallback(k, m, n, S, R)
    assert got.dtype == np.int32 and got.tolist() == TA["indices"][sample]


@pytest.mark.parametrize("k", [3, 5, 8, 16])
def test_shards_and_chunks_fold_to_the_same_keys(nn, oracle, k):
    """Feeding the reference set in pieces with global index bases (what shards and H2D chunks do)
    gives exactly the keys of one call; duplicates straddle the piece boundaries."""
    import torch
    from multicore_hw2_b200 import device
    S, R = cases.make("duplicated", 900 + k, k, 301, 20011)
    want = oracle.keys(S, R)
    dS, dR = torch.from_numpy(S).cuda(), torch.from_numpy(R).cuda()
    for pieces in (2, 3, 8):
        keys = device.new_keys(S.shape[0])
        for p in reversed(range(pieces)):  # order must not matter
            b, c = nn.shard_range(R.shape[0], pieces, p)
            if c:
                device.nearest_keys(dS, dR[b:b + c], keys, b)
        assert np.array_equal(keys.cpu().numpy().view(np.uint64), want)
        assert np.array_equal(device.keys_unpack(keys).cpu().numpy(), (want & 0xFFFFFFFF).astype(np.int32))


@pytest.mark.parametrize("variant", ["rreg", "rtma"])
@pytest.mark.parametrize("k,m,n", [(8, 8, 300007), (3, 7, 250001), (16, 13, 120000), (5, 1, 199999), (12, 2, 65536)])
def test_few_query_kernels_over_many_tiles(nn, oracle, variant, k, m, n):
    """The two few-query kernels on reference sets that span many tiles and every CTA of the
    persistent grid, with duplicated references (lowest index must win across tiles, warps and
    CTAs) and a ragged end; bit-exact keys against the oracle."""
    S, R = cases.make("duplicated", 5100 + k + m, k, m, n)
    assert np.array_equal(gpu_keys(nn, S, R, variant), oracle.keys(S, R))


@pytest.mark.parametrize("k", list(range(3, 17)))
def test_wide_query_tiles_on_adversarial_data(nn, oracle, k):
    """The query-register kernel with its WIDE tiles forced (the planner prefers narrow ones for the
    small fixture shapes) on the arithmetic- and tie-sensitive families: 8, 4 and the default number of
    queries per thread, enough queries to fill several tiles, bit-exact keys."""
    for kind in ("twins", "specials", "duplicated"):
        S, R = cases.make(kind, 7100 + k, k, 1100, 2999)
        want = oracle.keys(S, R)
        for q in (0, 4, 8):
            try:
                got = gpu_keys(nn, S, R, "qreg", q=q)
            except nn.NNError:
                assert (q == 8 and k > 8) or (q == 4 and k in (13, 15)), (k, q)   # tile width not built for this k
                continue
            assert np.array_equal(got, want), (kind, k, q)


@pytest.mark.parametrize("k", list(range(3, 17)))
def test_super_chunk_loop_on_adversarial_data(nn, oracle, k):
    """The query-register kernel's super-chunk loop form (used on splits of 64K+ references, for the k it
    is built for): forced with windows of 2, 3 and 8 chunks -- which straddle tiles, the end of a split
    and the ragged end of the set -- on the tie- and arithmetic-sensitive families, every tile width."""
    try:
        nn.set_option("variant", VARIANTS["qreg"])
        nn.set_option("qreg_super", 3)
        if "super=3" not in nn.describe_plan(k, 1100, 5003):
            pytest.skip(f"super-chunk form not built for k={k}")
        for kind, m, n in (("twins", 300, 2999), ("duplicated", 1100, 5003), ("specials", 300, 1001)):
            S, R = cases.make(kind, 7200 + k, k, m, n)
            want = oracle.keys(S, R)
            for sc in (2, 3, 8):
                nn.set_option("qreg_super", sc)
                nn.set_option("splits", 1 if sc == 3 else 0)   # one split: many tiles per CTA, ring re-used
                for q in (0, 2, 4, 8):
                    try:
                        got = gpu_keys(nn, S, R, "qreg", q=q)
                    except nn.NNError:
                        assert (q == 8 and k > 8) or (q == 4 and k in (13, 15)), (k, q)
                        continue
                    assert np.array_equal(got, want), (kind, k, sc, q)
    finally:
        nn.set_option("qreg_super", -1)
        nn.set_option("splits", 0)
        nn.set_option("variant", 0)


@pytest.mark.parametrize("variant", ["rreg", "rtma"])
@pytest.mark.parametrize("k,m,n", [(8, 10, 300007), (3, 11, 250001), (16, 12, 120000), (5, 3, 199999), (7, 9, 70001),
                                   (12, 4, 90001), (4, 2, 150001)])
def test_few_query_tail_passes_across_tiles(nn, oracle, variant, k, m, n):
    """Query counts whose last pass is 1..4 queries wide (the MQ = 2 and MQ = 4 instantiations) over
    reference sets spanning many tiles of every CTA, duplicates across tile boundaries, ragged end."""
    S, R = cases.make("duplicated", 7300 + k + m, k, m, n)
    assert np.array_equal(gpu_keys(nn, S, R, variant), oracle.keys(S, R))


@pytest.mark.parametrize("super_chunks", [-1, 3])
def test_nn_bench_cross_check_sweep(nn, super_chunks):
    """`nn_bench --sweep check`: every k, every kernel family, awkward sizes, 8-level quantised data
    (ties everywhere), each tuned kernel against the independent plain kernel on the device; once more
    with the query-register kernel's super-chunk loop forced (3 chunks per update)."""
    import subprocess
    exe = os.path.join(os.path.dirname(nn.LIB_PATH), "nn_bench")
    if not os.path.exists(exe):
        pytest.skip("nn_bench not built")
    out = subprocess.run([exe, "--opt", f"qreg_super={super_chunks}", "--sweep", "check"], capture_output=True,
                         text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    rows = [json.loads(l) for l in out.stdout.splitlines() if l.startswith('{"op":"nearest_keys"')]
    assert len(rows) >= 14 * 6 and all(r["mismatch_vs_plain"] == 0 for r in rows), \
        [(r["k"], r["m"], r["plan"][:20]) for r in rows if r["mismatch_vs_plain"] != 0]


def test_nn_bench_fuzz_sweep(nn):
    """`nn_bench --sweep fuzz`: 240 pseudo-random shapes (every k, query counts on both sides of every
    kernel-family boundary, reference counts from 1 to 420k, a third on an 8-level grid) with the
    planner in charge, alternating between the building blocks and the one-launch search, each
    against the plain kernel on the device."""
    import subprocess
    exe = os.path.join(os.path.dirname(nn.LIB_PATH), "nn_bench")
    if not os.path.exists(exe):
        pytest.skip("nn_bench not built")
    out = subprocess.run([exe, "--sweep", "fuzz"], capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stderr[-2000:]
    rows = [json.loads(l) for l in out.stdout.splitlines() if l.startswith('{"op":')]
    assert len(rows) == 240 and all(r["mismatch_vs_plain"] == 0 for r in rows), \
        [(r["k"], r["m"], r["n"], r["quant"], r["op"], r["plan"][:24]) for r in rows if r["mismatch_vs_plain"] != 0]


@pytest.mark.parametrize("k", list(range(3, 17)))
def test_phased_query_register_kernel_layouts(nn, oracle, k):
    """nn_qflex_kernel: 128 threads = query groups x reference phases.  Query counts that give very
    different layouts (many phases / few, several query tiles, forced queries per thread), reference
    counts with ragged rounds and several tiles per CTA, duplicated references across phases, tiles and
    CTAs (lowest index must win); bit-exact keys against the oracle."""
    for m, n, q in [(9, 70001, 0), (31, 40003, 2), (100, 50021, 0), (100, 3, 4), (130, 33333, 0), (257, 20011, 2),
                    (300, 9973, 4), (17, 129, 0)]:
        if q == 4 and k in (13, 15):
            continue  # 4 queries per thread do not fit the register budget there (QregDefault)
        S, R = cases.make("duplicated", 6100 + k + m, k, m, n)
        assert np.array_equal(gpu_keys(nn, S, R, "qflex", q=q), oracle.keys(S, R)), (k, m, n, q)
    S, R = cases.make("twins", 6200 + k, k, 100, 3000)
    assert np.array_equal(gpu_keys(nn, S, R, "qflex"), oracle.keys(S, R))
    S, R = cases.make("specials", 6300 + k, k, 77, 1000)
    assert np.array_equal(gpu_keys(nn, S, R, "qflex"), oracle.keys(S, R))


@pytest.mark.parametrize("k,n", [(3, 10007), (7, 4096), (8, 5001), (16, 3000), (13, 1)])
def test_repack_soa_matches_mat_inv(nn, oracle, k, n):
    import torch
    from multicore_hw2_b200 import device
    R = np.random.default_rng(k * 31 + n).random((n, k), dtype=np.float32)
    got = device.repack_soa(torch.from_numpy(R).cuda()).cpu().numpy()
    assert np.array_equal(got.view(np.uint32), oracle.repack_soa(R).view(np.uint32))


@pytest.mark.parametrize("threads", [0, 1, 3])
def test_pageable_inputs_through_the_staging_threads(nn, oracle, threads):
    """The reference's harness passes malloc()ed arrays (generator.h:37, 44).  For pageable inputs the
    host entry stages the reference chunks through pinned buffers on several host threads; small
    chunks force many of them, with duplicated references straddling the chunk boundaries."""
    S, R = cases.make("duplicated", 8800 + threads, 5, 77, 60013)
    nn.set_option("h2d_chunk_bytes", 1)       # clamps to the minimum chunk of 4096 references
    nn.set_option("stage_threads", threads)
    try:
        for _ in range(2):                    # second call reuses the staging buffers
            assert np.array_equal(nn.search_host(S, R, num_gpus=1), oracle.v0(S, R, threads=0))
    finally:
        nn.set_option("h2d_chunk_bytes", 16 << 20)
        nn.set_option("stage_threads", -1)


def test_resident_index_build_once_query_many(nn, oracle):
    """nn_b200_index_*: the references stay in HBM, several query batches of different sizes (each
    kernel family: 1, 7, 300 queries) are answered against them; identical to v0 and to the
    one-shot host entry.  Also on every GPU count available."""
    import torch
    S, R = cases.make("duplicated", 4711, 6, 300, 70001)
    want = oracle.v0(S, R, threads=0)
    for gpus in sorted({1, min(2, torch.cuda.device_count()), torch.cuda.device_count()}):
        for graph in (1, 0):  # small batches on one GPU replay a captured CUDA graph; 0 = plain enqueues
            nn.set_option("index_graph", graph)
            with nn.Index(R, num_gpus=gpus) as ix:
                assert (ix.k, ix.n, ix.gpus) == (6, 70001, gpus)
                for lo, hi in [(0, 1), (1, 8), (0, 300), (293, 300), (0, 1), (1, 8)]:  # repeats replay the graph
                    assert np.array_equal(ix.search(S[lo:hi]), want[lo:hi]), (gpus, graph, lo, hi)
                assert ix.search(S[:0]).size == 0
    nn.set_option("index_graph", 1)
    with nn.Index(np.zeros((0, 3), np.float32), k=3) as ix:   # empty index: v0's start state, index 0
        assert ix.search(np.ones((4, 3), np.float32)).tolist() == [0] * 4
    with pytest.raises(nn.NNError):
        nn.Index(np.zeros((5, 2), np.float32))


def test_edge_shapes(nn, oracle):
    import torch
    from multicore_hw2_b200 import device
    # n = 0: every query keeps v0's start state -> index 0 (core.cu:39-40)
    out = nn.search_host(np.zeros((5, 3), np.float32), np.zeros((0, 3), np.float32), k=3)
    assert out.tolist() == [0] * 5
    assert nn.search_host(np.zeros((0, 3), np.float32), np.zeros((4, 3), np.float32), k=3).size == 0
    for k, m, n in [(3, 1, 1), (16, 1, 1), (4, 2, 3), (9, 1, 5), (16, 129, 2), (3, 1, 2)]:
        S, R = cases.make("quantized", k + m + n, k, m, n)
        assert np.array_equal(nn.search_host(S, R), oracle.v0(S, R)), (k, m, n)
    with pytest.raises(nn.NNError):
        nn.search_host(np.zeros((1, 2), np.float32), np.zeros((1, 2), np.float32), k=2)
    with pytest.raises(nn.NNError):
        nn.search_host(np.zeros((1, 17), np.float32), np.zeros((1, 17), np.float32), k=17)
    # misaligned reference pointer is refused, not mis-read
    S = torch.zeros((4, 3), device="cuda")
    R = torch.zeros((9, 3), device="cuda")
    with pytest.raises(nn.NNError):
        device.nearest_keys(S, R.view(-1)[1:25].view(8, 3), device.new_keys(4))


# ---- BASELINE.json sizes ---------------------------------------------------------------------------

def _device_uniform(seed, rows, k):
    import torch
    g = torch.Generator(device="cuda")
    g.manual_seed(seed)
    return torch.rand((rows, k), generator=g, device="cuda", dtype=torch.float32)


def _check_subset_against_oracle(nn, oracle, dS, dR, keys, rows):
    S = dS[rows].cpu().numpy()
    R = dR.cpu().numpy()
    want = oracle.keys(S, R)
    got = keys[rows].cpu().numpy().view(np.uint64)
    assert np.array_equal(got, want)


def test_config1_is_ta_sample_6(nn, oracle):
    """BASELINE config 1 (k=3, m=1024, n=65536) = TA sample 6 = results.csv line 13."""
    S, R = oracle.ta_sample(6)
    assert nn.cudaCallback(3, 1024, 65536, S, R).tolist() == TA["indices"][6]
    assert (gpu_keys(nn, S, R) & 0xFFFFFFFF).tolist() == TA["indices"][6]


def test_config2_full_size(nn, oracle):
    """k=16, m=4096, n=2^20: oracle on 96 queries; tuned kernel == plain kernel on all 4096."""
    import torch
    from multicore_hw2_b200 import device
    dS, dR = _device_uniform(1002, 4096, 16), _device_uniform(2002, 1 << 20, 16)
    keys = device.nearest_keys(dS, dR, device.new_keys(4096))
    nn.set_option("variant", 3)
    plain = device.nearest_keys(dS, dR, device.new_keys(4096))
    nn.set_option("variant", 0)
    assert torch.equal(keys, plain)
    rows = torch.arange(0, 4096, 43, device="cuda")
    _check_subset_against_oracle(nn, oracle, dS, dR, keys, rows)


def test_config3_full_size(nn, oracle):
    """k=8, m=8, n=2^26: all 8 queries against the oracle (all host threads)."""
    from multicore_hw2_b200 import device
    dS, dR = _device_uniform(1003, 8, 8), _device_uniform(2003, 1 << 26, 8)
    keys = device.nearest_keys(dS, dR, device.new_keys(8))
    _check_subset_against_oracle(nn, oracle, dS, dR, keys, list(range(8)))
    del dR


def test_config5_ties_properties(nn, oracle):
    """k=3, m=n=2^20 on an 8-bit grid (every distance is tied many times over):
    * each query that is itself a reference must return the LOWEST index holding that point;
    * appending a copy of the reference set (indices n..2n-1) must not change any answer;
    * a sample of queries matches the oracle."""
    import torch
    from multicore_hw2_b200 import device
    n = 1 << 20
    dR = torch.floor(_device_uniform(2005, n, 3) * 256) / 256
    dS = dR[torch.randperm(n, device="cuda", generator=torch.Generator(device="cuda").manual_seed(5))].contiguous()
    keys = device.nearest_keys(dS, dR, device.new_keys(n))
    idx = device.keys_unpack(keys).long()
    assert int((keys >> 32).max()) == 0  # distance 0 everywhere
    assert torch.equal(dR[idx], dS)
    # lowest index among identical points: first occurrence of each distinct row
    code = (dR * 256).long()
    code = (code[:, 0] << 16) | (code[:, 1] << 8) | code[:, 2]
    first = torch.full((1 << 24,), n, dtype=torch.long, device="cuda")
    first.scatter_reduce_(0, code, torch.arange(n, device="cuda"), reduce="amin")
    scode = (dS * 256).long()
    scode = (scode[:, 0] << 16) | (scode[:, 1] << 8) | scode[:, 2]
    assert torch.equal(idx, first[scode])
    keys2 = device.nearest_keys(dS, torch.cat([dR, dR]), device.new_keys(n))
    assert torch.equal(keys2, keys)
    rows = torch.arange(0, n, 16411, device="cuda")
    _check_subset_against_oracle(nn, oracle, dS, dR, keys, rows)


def test_config4_scaled_and_sharded(nn, oracle):
    """k=16, m=65536 against a 2^21-reference slice of config 4 (the full 2^24 runs in bench.py),
    fed as 8 shards with global index bases; oracle on 48 queries."""
    import torch
    from multicore_hw2_b200 import device
    n = 1 << 21
    dS, dR = _device_uniform(1004, 65536, 16), _device_uniform(2004, n, 16)
    keys = device.new_keys(65536)
    for p in range(8):
        b, c = nn.shard_range(n, 8, p)
        device.nearest_keys(dS, dR[b:b + c], keys, b)
    whole = device.nearest_keys(dS, dR, device.new_keys(65536))
    assert torch.equal(keys, whole)
    rows = torch.arange(0, 65536, 1400, device="cuda")
    _check_subset_against_oracle(nn, oracle, dS, dR, keys, rows)


def test_config4_full_size(nn, oracle):
    """BASELINE config 4 at FULL size (k=16, m=65536, n=2^24; the bench default): the one-launch search
    against the oracle on 24 queries over all 2^24 references, and against the same set folded as 8 shards
    with global index bases (what 8 GPUs compute) on all 65536 queries."""
    import torch
    from multicore_hw2_b200 import device
    m, n = 65536, 1 << 24
    dS, dR = _device_uniform(1004, m, 16), _device_uniform(2004, n, 16)
    ws = device.Workspace(m)
    idx = device.search(dS, dR, ws)
    keys = device.new_keys(m)
    for p in range(8):
        b, c = nn.shard_range(n, 8, p)
        device.nearest_keys(dS, dR[b:b + c], keys, b)
    assert torch.equal(device.keys_unpack(keys), idx)
    rows = torch.arange(0, m, 2801, device="cuda")
    _check_subset_against_oracle(nn, oracle, dS, dR, keys, rows)
    del dR


def test_multi_gpu_host_entry(nn, oracle):
    """v8's job: the host entry shards the references over all visible GPUs and merges the keys --
    by system-scope atomicMin into GPU 0's key array from inside every GPU's search kernel (default)
    or with ncclAllReduce(min, uint64).  Needs >= 2 GPUs (gpurun --gpus 2); duplicates straddle the
    shard boundaries so the lowest-index rule is exercised across devices."""
    import torch
    g = torch.cuda.device_count()
    if g < 2:
        pytest.skip("needs at least 2 GPUs")
    for kind, k, m, n in [("duplicated", 16, 300, 40001), ("twins", 8, 200, 30011), ("quantized", 3, 1000, 5),
                          ("uniform", 3, 1, 1)]:
        S, R = cases.make(kind, 7000 + k, k, m, n)
        want = oracle.v0(S, R, threads=0)
        for gpus in sorted({2, g}):
            for p2p in (1, 0):  # fold into GPU 0's keys over NVLink inside the kernels / NCCL all-reduce
                nn.set_option("p2p_merge", p2p)
                assert np.array_equal(nn.search_host(S, R, num_gpus=gpus), want), (kind, gpus, p2p)
    nn.set_option("p2p_merge", 1)
    # same through the reference's own entry point, all GPUs
    S, R = oracle.ta_sample(7)
    assert nn.cudaCallback(16, 1024, 65536, S, R).tolist() == TA["indices"][7]


def test_peer_merge_between_processes(nn):
    """One process per GPU (torchrun): every rank's search kernel folds into rank 0's key array over
    NVLink peer memory (CUDA IPC) and rank 0's last CTA stores the merged indices
    (multicore_hw2_b200.sharded.PeerMerge, include/nn_b200.h 2c).  18 searches in a row -- every kernel
    family, empty shards, alternating key buffers -- each compared with the CPU oracle on rank 0."""
    import subprocess
    import sys
    import torch
    g = torch.cuda.device_count()
    if g < 2:
        pytest.skip("needs at least 2 GPUs")
    worker = os.path.join(os.path.dirname(os.path.abspath(__file__)), "peer_merge_worker.py")
    for world in sorted({2, min(g, 4), g}):
        out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                              "--master-addr", "127.0.0.1", "--master-port", str(29600 + world), worker],
                             capture_output=True, text=True, timeout=600)
        assert out.returncode == 0, (world, out.stdout[-1500:], out.stderr[-1500:])
        assert "rank 0: peer merge ok" in out.stdout


def test_device_blocks_on_every_gpu(nn, oracle):
    """The device-resident entry points on cuda:1.. after cuda:0 (per-device shared-memory opt-in of
    every kernel instantiation: the MQ = 2/4 tails of the reference-stream kernel and the phased kernel
    were once configured for the first device only).  Needs >= 2 GPUs."""
    import torch
    from multicore_hw2_b200 import device
    if torch.cuda.device_count() < 2:
        pytest.skip("needs at least 2 GPUs")
    for m, variant in [(11, "rtma"), (10, "rtma"), (100, "qflex"), (300, "qreg"), (3, "rreg")]:
        S, R = cases.make("duplicated", 7400 + m, 8, m, 50021)
        want = oracle.keys(S, R)
        nn.set_option("variant", VARIANTS[variant])
        try:
            for d in range(torch.cuda.device_count()):
                with torch.cuda.device(d):
                    dS, dR = torch.from_numpy(S).cuda(d), torch.from_numpy(R).cuda(d)
                    keys = device.nearest_keys(dS, dR, device.new_keys(m, dS.device))
                    ws = device.Workspace(m, dS.device)
                    idx = device.search(dS, dR, ws)
                    torch.cuda.synchronize(d)
                    assert np.array_equal(keys.cpu().numpy().view(np.uint64), want), (m, variant, d)
                    assert np.array_equal(idx.cpu().numpy(), (want & 0xFFFFFFFF).astype(np.int32)), (m, variant, d)
        finally:
            nn.set_option("variant", 0)


def test_reference_harness_runs_unmodified_against_the_library(nn):
    """oracle/_ref/harness_main = the reference's TA harness (main.cu, generator.h, utils.h, compiled
    unmodified from /root/reference by oracle/Makefile) + the reference's own v0 as Callback1 + THIS
    library as Callback10 through integration/core.h.  Its own checker (main.cu:84-97) must report
    0 errors on all eight samples, i.e. the output of the reference's screen.log for Callback10."""
    import subprocess
    exe = os.path.join(os.path.dirname(GOLD), "..", "oracle", "_ref", "harness_main")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/harness_main not built (needs /root/reference at build time)")
    env = dict(os.environ, NN_B200_STATIC_WARMUP="1")   # the counterpart of the reference's static WarmUP (core.cu:1274)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    txt = out.stdout
    assert "on running CALLBACK10" in txt
    errs = [l for l in txt.splitlines() if l.startswith("errors/total")]
    assert len(errs) == 8 and all(l.split(":")[1].strip().startswith("0/") for l in errs), txt[-1500:]
    shapes = [l.split(",")[1:4] for l in txt.splitlines() if l.startswith("Callback2,")]
    assert [[int(x) for x in s] for s in shapes] == TA["samples"]
    # with the static warm-up the FIRST timed sample no longer pays for context creation and code loading
    first = [l for l in txt.splitlines() if l.startswith("Callback2,")][0]
    assert float(first.split(",")[4].strip().rstrip("ms")) < 50.0, first
