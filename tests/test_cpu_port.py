"""Pins the oracle port (oracle/port, plain C, independent of /root/reference) on the reference's goldens, on the
committed outputs of the compiled reference, and — where oracle/_ref is present — on the compiled reference live."""
import os

import numpy as np
import pytest

import portso
import refso
from golden import proto_cross_terms as proto

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")
needs_ref = pytest.mark.skipif(not refso.available(), reason="compiled reference (oracle/_ref) not present")


@pytest.fixture(scope="module")
def G():
    return np.load(os.path.join(GOLD, "golden_4g9s.npz"))


def test_port_expansion_matches_ref_spf_and_reference_outputs(G):
    """src/pdb2spf.c restated: ref_spf golden (V, D at its print precision) and the compiled reference's coefficients"""
    q, L = G["qvals"], int(G["L"])
    ff = G["rec_ff"].copy()
    ff[:, 2] = ff[:, 2] * G["rec_sa"]          # h2o zero_ff * SASA fraction
    coef = portso.expand(G["rec_xyz"], ff, q, L)
    assert np.array_equal(coef, G["rec_coef"])  # same arithmetic, same bits
    ref = G["ref_spf"]
    for c in range(2):
        mine, r = coef[c], ref[:, :, 2 * c:2 * c + 2]
        nz = (np.abs(mine) > 0) & (np.abs(r) > 0)
        assert np.max(np.abs((r - mine)[nz] / r[nz])) < 1e-3
    assert np.array_equal(portso.mkarray(0.0, 0.5, 50), q)


def test_port_scores_reproduce_ref_chi_golden(G):
    """tests/saxs_test.c score_conformations: 52 rows vs ref_chi (abs 1e-3) and vs the compiled reference (1e-9)"""
    q, L = G["qvals"], int(G["L"])
    a, scal = portso.opt_params(G["exp_q"], G["exp_in"], G["exp_err"], q, float(G["scal"][0]))
    assert np.array_equal(a, G["a"]) and np.array_equal(scal, G["scal"])
    s, c1, c2, X = portso.scores(G["z40_index"], G["rec_coef"], G["lig_coef"], a, scal, q, [40.0], L, want_cross=True)
    rc = G["ref_chi"]
    assert np.max(np.abs(s - rc[:, 1])) < 1e-3 and np.max(np.abs(c1 - rc[:, 2])) < 1e-3 and np.max(np.abs(c2 - rc[:, 3])) < 1e-3
    assert np.max(np.abs(s / G["z40_scores"] - 1)) < 1e-9
    assert np.max(np.abs(c1 / G["z40_c1"] - 1)) < 1e-7
    # the reference-structured cross terms agree with the factorised numpy restatement used for the CUDA path
    FC = np.load(os.path.join(GOLD, "fit_cases.npz"))
    assert np.max(np.abs(X - FC["X52"]) / np.abs(FC["X52"]).max(axis=2, keepdims=True)) < 1e-10


def test_port_fit_bitwise_on_fixture(G):
    FC = np.load(os.path.join(GOLD, "fit_cases.npz"))
    X = np.concatenate([FC["X52"], proto.perturbed_family(FC["X52"], 4000, 1)])
    want = np.concatenate([FC["fit52"], FC["fit_family"]])
    assert np.array_equal(portso.fit(X, G["a"], G["scal"], G["qvals"]), want)


def test_port_edge_cases(G):
    """rows off the z table keep incoming values; duplicates agree; empty list"""
    q, L = G["qvals"], int(G["L"])
    nb, N = L + 1, 2 * L + 1
    idx = np.concatenate([G["z40_index"][:3], G["z40_index"][:1], [5 * nb * nb * N ** 3 + 3, -1]]).astype(np.int64)
    init = (np.full(6, 7.0), np.full(6, 8.0), np.full(6, 9.0))
    s, c1, c2 = portso.scores(idx, G["rec_coef"], G["lig_coef"], G["a"], G["scal"], q, [40.0], L, init=init)
    assert s[3] == s[0] and c1[3] == c1[0]
    assert (s[4], c1[4], c2[4]) == (7.0, 8.0, 9.0) and (s[5], c1[5], c2[5]) == (7.0, 8.0, 9.0)
    assert len(portso.scores(np.zeros(0, dtype=np.int64), G["rec_coef"], G["lig_coef"], G["a"], G["scal"], q, [40.0], L)[0]) == 0


@needs_ref
def test_port_tables_match_compiled_reference():
    M_PI = 3.14159265358
    rng = np.random.default_rng(1)
    for x in rng.uniform(0, 45, 300):
        for l in (0, 3, 15, 30):
            assert portso.sbessel(l, x) == refso.sbessel(l, x)
    for L in (7, 15):
        for k in (0, 2, L):
            assert np.array_equal(portso.wigner_d(L, k * (M_PI / L)), refso.wigner_d(L, k * (M_PI / L)))
    L = 15
    ds = portso.dsymb(L)
    dsymb, _, _ = proto.reference_tables(L, np.array([0.1]), [1.0])
    for l in range(L + 1):
        for m in range(-l, l + 1):
            assert np.max(np.abs(ds[l * (l + 1) + m] - dsymb[l, m + L])) < 1e-14


@needs_ref
def test_port_scores_match_compiled_reference_live(G):
    """a few fresh grid points in two z steps, incl. one cell with >= 30 rows (the reference's FFT branch)"""
    q, L = G["qvals"], int(G["L"])
    nb, N = L + 1, 2 * L + 1
    rng = np.random.default_rng(3)
    rows = []
    for _ in range(34):
        rows.append((((((0 * nb + 6) * nb + 9) * N + rng.integers(0, N)) * N + rng.integers(0, N)) * N + rng.integers(0, N)))
    for _ in range(4):
        rows.append((((((1 * nb + rng.integers(1, nb - 1)) * nb + 9) * N + rng.integers(0, N)) * N + rng.integers(0, N)) * N + rng.integers(0, N)))
    idx = np.array(rows, dtype=np.int64)
    zv = [25.0, 61.5]
    want = refso.scores(idx.astype(np.int32), G["rec_coef"], G["lig_coef"], G["a"], G["scal"], q, zv, L)
    got = portso.scores(idx, G["rec_coef"], G["lig_coef"], G["a"], G["scal"], q, zv, L)
    assert np.max(np.abs(got[0] / want[0] - 1)) < 1e-8
    assert np.max(np.abs(got[1] / want[1] - 1)) < 1e-6
    assert np.max(np.abs(got[2] - want[2])) < 4e-6
