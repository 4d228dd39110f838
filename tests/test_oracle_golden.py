"""Pins the CPU oracle (oracle/nn_oracle.c) before anything is compared against it:

* against the reference's own golden file results.csv (tests/golden/ta_results.json), whose odd
  lines are v0's nearest indices for the eight TA samples of main.cu:28-39 under srand(1000);
* against outputs of the reference's own v0::cudaCallback compiled from /root/reference
  (tests/golden/ref_v0_cases.json, written by tests/golden/make_golden.py);
* and, when oracle/_ref is present, against that library live.
All CPU, a few seconds."""
import json
import os

import numpy as np
import pytest

import cases

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TA = json.load(open(os.path.join(GOLD, "ta_results.json")))
REF = json.load(open(os.path.join(GOLD, "ref_v0_cases.json")))


@pytest.mark.parametrize("sample", range(8))
def test_ta_samples_match_results_csv(oracle, sample):
    k, m, n = oracle.ta_shape(sample)
    assert [k, m, n] == TA["samples"][sample]
    S, R = oracle.ta_sample(sample)
    threads = 1 if m * n < 1 << 22 else 0
    idx, dist = oracle.v0(S, R, k, threads=threads, want_dist=True)
    assert idx.tolist() == TA["indices"][sample]
    if sample >= 2:  # results.csv distance lines 2 and 4 do not reproduce
        np.testing.assert_allclose(np.sqrt(dist), np.array(TA["distances"][sample]), atol=5.1e-4)


def test_mt_oracle_is_bit_identical_to_serial(oracle):
    S, R = cases.make("twins", 11, 7, 257, 3001)
    a, da = oracle.v0(S, R, threads=1, want_dist=True)
    b, db = oracle.v0(S, R, threads=0, want_dist=True)
    assert np.array_equal(a, b) and np.array_equal(da.view(np.uint32), db.view(np.uint32))
    keys = oracle.keys(S, R)
    assert np.array_equal((keys & 0xFFFFFFFF).astype(np.int32), a)
    assert np.array_equal((keys >> 32).astype(np.uint32), da.view(np.uint32))


@pytest.mark.parametrize("case", REF["cases"], ids=lambda c: f"{c['kind']}-k{c['k']}-m{c['m']}-n{c['n']}")
def test_restatement_matches_reference_v0_fixture(oracle, case):
    S, R = cases.make(case["kind"], case["seed"], case["k"], case["m"], case["n"])
    assert cases.checksum(S, R) == case["crc"], "input generator drifted; regenerate tests/golden"
    assert oracle.v0(S, R, case["k"]).tolist() == case["indices"]


def test_restatement_matches_reference_v0_live(oracle):
    if not oracle.ref_available():
        pytest.skip("oracle/_ref not built (no /root/reference on this machine)")
    for kind, seed, k, m, n in [("twins", 77, 16, 128, 4096), ("quantized", 78, 3, 512, 2048),
                                ("specials", 79, 8, 64, 1000), ("uniform", 80, 13, 50, 3333)]:
        S, R = cases.make(kind, seed, k, m, n)
        ref, _ = oracle.ref_v0(S, R, k)
        assert np.array_equal(oracle.v0(S, R, k), ref)
        ref_mt, used = oracle.ref_v0(S, R, k, threads=0)
        assert np.array_equal(ref_mt, ref) and used >= 1


def test_twins_family_is_arithmetic_sensitive(oracle):
    """The `twins` inputs must separate v0's arithmetic from an FMA-contracted one; otherwise a
    kernel with the wrong arithmetic could pass the parity tests by luck."""
    S, R = cases.make("twins", 408, 8, 300, 2508)
    want = oracle.v0(S, R)
    d = (S[:, None, :] - R[None, :, :]).astype(np.float32)
    acc = np.zeros(d.shape[:2], np.float32)
    for j in range(d.shape[2]):  # fma(d, d, acc): exact product, one rounding
        dd = d[:, :, j].astype(np.float64)
        acc = (dd * dd + acc.astype(np.float64)).astype(np.float32)
    fused = np.argmin(acc, axis=1)
    assert (fused != want).sum() > 30


def test_repack_soa_oracle(oracle):
    R = np.arange(5 * 3, dtype=np.float32).reshape(5, 3)
    assert np.array_equal(oracle.repack_soa(R), R.T)
