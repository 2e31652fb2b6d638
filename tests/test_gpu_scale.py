"""Larger and differently-shaped GPU parity cases: the full real pose list (70 000 rows, 23 z steps) against stored
outputs of the compiled reference, BASELINE config 4's shape (L = 30, Q = 100), and size-independent properties on the
synthetic workload generator used by bench.py."""
import os

import numpy as np
import pytest

from libfmftsaxs_b200 import capi
from libfmftsaxs_b200 import workload as wl
import parity
import refso

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")
pytestmark = pytest.mark.gpu
TOL = 1e-6


def names(a):
    return [x.decode() for x in a]


def test_real_list_70k_rows_against_reference():
    """tests/data/euler_coords.000.00 complete: 70 000 rows, z = 20..42 on the 1..80 table, 42 246 distinct grid
    points, cells from 1 to 383 rows (both DFT branches of the reference)"""
    G = np.load(os.path.join(GOLD, "golden_4g9s.npz"))
    R = np.load(os.path.join(GOLD, "golden_real70k.npz"))
    q, L = G["qvals"], int(G["L"])
    s, c1, c2 = capi.scores(R["index"], G["rec_coef"], G["lig_coef"], G["a"], G["scal"], q, R["zvals"], L)
    parity.check("real list, 70 000 rows", (s, c1, c2), (R["scores"], R["c1"], R["c2"]),
                 sens=(R["sens_scores"], R["sens_c1"], R["sens_c2"]))
    # text-level contract of the output file: three decimals
    assert np.mean(np.round(s, 3) == np.round(R["scores"], 3)) > 0.999


def test_config4_shape_L30_Q100():
    F = np.load(os.path.join(GOLD, "golden_l30.npz"))
    L, q = int(F["L"]), F["qvals"]
    A, _, _ = capi.expand(wl.MAP_PATH, F["rec_xyz"], names(F["rec_res"]), names(F["rec_atm"]), F["rec_radius"], q, L,
                          sa=F["rec_sa"], water_mode=1)
    B, _, _ = capi.expand(wl.MAP_PATH, F["lig_xyz"], names(F["lig_res"]), names(F["lig_atm"]), F["lig_radius"], q, L,
                          sa=F["lig_sa"], water_mode=1)
    sa = np.abs(F["coefA_sample"]).max()
    assert np.max(np.abs(A[:, ::25, ::97] - F["coefA_sample"])) / sa < 1e-9
    assert np.max(np.abs(B[:, ::25, ::97] - F["coefB_sample"])) / np.abs(F["coefB_sample"]).max() < 1e-9
    for idx in (F["index"], F["index"].astype(np.int64)):      # 32-bit and 64-bit entry points
        s, c1, c2 = capi.scores(idx, A, B, F["a"], F["scal"], q, F["zvals"], L)
        parity.check("L = 30, Q = 100, 6 poses", (s, c1, c2), (F["scores"], F["c1"], F["c2"]))
    # beyond 9 z steps the flat index needs 64 bits at L = 30 (SURVEY header note 6): z digit 40 of a 64-step table
    nb, N = L + 1, 2 * L + 1
    big = F["index"].astype(np.int64) % (nb * nb * N ** 3) + 40 * (nb * nb * N ** 3)
    assert big.max() > 2 ** 31
    zv = np.concatenate([np.full(40, 50.0), [F["zvals"][0]], np.full(23, 60.0)])
    first = (F["index"].astype(np.int64) // (nb * nb * N ** 3)) == 0
    s, c1, c2 = capi.scores(big[first], A, B, F["a"], F["scal"], q, zv, L)
    assert np.max(np.abs(s / F["scores"][first] - 1)) < TOL


def _bench_fixture_case(fname, tag):
    """the benchmarked workload against the compiled reference: (i) the scoring path on the reference's own coefficient
    tables, (ii) K1 on the device against those tables, (iii) the whole path from the atoms"""
    F = np.load(os.path.join(GOLD, fname))
    L, q = int(F["L"]), F["qvals"]
    want = (F["scores"], F["c1"], F["c2"])
    sens = (F["sens_scores"], F["sens_c1"], F["sens_c2"])
    got = capi.scores(F["index"], F["coefA"], F["coefB"], F["a"], F["scal"], q, F["zvals"], L)
    parity.check(tag + ", reference coefficients", got, want, sens=sens)
    got64 = capi.scores(F["index"].astype(np.int64), F["coefA"], F["coefB"], F["a"], F["scal"], q, F["zvals"], L)
    assert all(np.array_equal(x, y) for x, y in zip(got, got64))
    w = wl.make(str(F["workload"]), nrot=4, nz=1)
    A, _, _ = capi.expand(wl.MAP_PATH, w["rec"]["xyz"], w["rec"]["res"], w["rec"]["atm"], w["rec"]["radius"], q, L,
                          sa=w["rec"]["sa"], water_mode=1)
    B, _, _ = capi.expand(wl.MAP_PATH, w["lig"]["xyz"], w["lig"]["res"], w["lig"]["atm"], w["lig"]["radius"], q, L,
                          sa=w["lig"]["sa"], water_mode=1)
    for mine, ref in ((A, F["coefA"]), (B, F["coefB"])):
        scale = np.abs(ref).max(axis=(2, 3), keepdims=True)
        assert np.max(np.abs(mine - ref) / scale) < 1e-9
    got = capi.scores(F["index"], A, B, F["a"], F["scal"], q, F["zvals"], L)
    parity.check(tag + ", device coefficients", got, want, sens=sens)
    return F


def test_bench_workload_cfg3_slabs_against_reference():
    """BASELINE config 3 as bench.py builds it (3000 + 1500 atoms, L = 15, Q = 50): all 11 931 poses of 4 complete
    (z, beta2) slabs of the 4.48 M-pose list; fixture by tests/golden/make_golden_bench.py cfg3"""
    F = _bench_fixture_case("golden_cfg3_slabs.npz", "cfg3, 4 slabs")
    assert len(F["index"]) > 10000


def test_bench_workload_cfg4_cells_against_reference():
    """BASELINE config 4 (20 000 + 5 000 atoms, L = 30, Q = 100): all poses of 4 complete cells of the 8.96 M-pose
    list; fixture by tests/golden/make_golden_bench.py cfg4"""
    F = _bench_fixture_case("golden_cfg4_cells.npz", "cfg4, 4 cells")
    assert len(np.unique(F["index"].astype(np.int64) // 61 ** 3)) == 4


def test_properties_on_bench_workload():
    """size-independent properties at the bench's shape: order invariance, duplicates, idempotence, shard union"""
    w = wl.make("cfg3_3k+1.5k_L15_Q50_70kx64z", nrot=3000, nz=6)
    q, L = w["qvals"], w["L"]
    A, _, _ = capi.expand(wl.MAP_PATH, w["rec"]["xyz"], w["rec"]["res"], w["rec"]["atm"], w["rec"]["radius"], q, L,
                          sa=w["rec"]["sa"], water_mode=1)
    B, _, _ = capi.expand(wl.MAP_PATH, w["lig"]["xyz"], w["lig"]["res"], w["lig"]["atm"], w["lig"]["radius"], q, L,
                          sa=w["lig"]["sa"], water_mode=1)
    eq, ei, ee = wl.experimental_curve(A, B, q)
    a, scal = capi.opt_params(eq, ei, ee, q, wl.mean_radius(w["rec"], w["lig"]))
    idx = w["index"]
    plan = capi.Plan(L, q)
    plan.set_molecules(A, B)
    plan.set_experiment(a, scal[1], scal[2])
    plan.set_translations(w["zvals"])
    base = plan.score(idx)
    assert all(np.isfinite(x).all() for x in base[1:])
    assert (base[1] >= 0.96).all() and (base[1] <= 1.04).all() and (base[2] >= -2).all() and (base[2] <= 4).all()
    again = plan.score(idx)
    assert all(np.array_equal(x, y, equal_nan=True) for x, y in zip(base, again))      # deterministic / idempotent
    perm = np.random.default_rng(0).permutation(len(idx))
    shuf = plan.score(idx[perm])
    assert all(np.array_equal(x[perm], y, equal_nan=True) for x, y in zip(base, shuf))  # order invariance, bit for bit
    dup = plan.score(np.concatenate([idx, idx[:1000]]))
    assert all(np.array_equal(x, y[:len(idx)], equal_nan=True) and np.array_equal(x[:1000], y[len(idx):], equal_nan=True)
               for x, y in zip(base, dup))                                                # duplicates share one fit
    # z shards: the union of per-range calls equals the whole call (the multi-GPU decomposition)
    parts = [np.full(len(idx), -7.0) for _ in range(3)]
    for lo, hi in ((0, 2), (2, 3), (3, 6)):
        got = plan.score(idx, z_lo=lo, z_hi=hi, init=tuple(parts))
        parts = list(got)
    assert all(np.array_equal(x, y, equal_nan=True) for x, y in zip(base, parts))
    # the flat-API path (sxs_compute_saxs_scores) gives the same bits as the plan path
    flat = capi.scores(idx, A, B, a, scal, q, w["zvals"], L)
    assert all(np.array_equal(x, y, equal_nan=True) for x, y in zip(base, flat))
    plan.close()
    # tight workspace budgets: 2 z steps per translated group, ~1000 points per cross-term chunk -> same bits
    old = {k: os.environ.get(k) for k in ("SXS_CUDA_ST_GB", "SXS_CUDA_X_GB")}
    os.environ["SXS_CUDA_ST_GB"], os.environ["SXS_CUDA_X_GB"] = "0.4", "0.002"
    try:
        small = capi.Plan(L, q)
        small.set_molecules(A, B)
        small.set_experiment(a, scal[1], scal[2])
        small.set_translations(w["zvals"])
        tight = small.score(idx)
        assert small.stats()["groups"] >= 3 and small.stats()["launches"] > 30
        small.close()
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    assert all(np.array_equal(x, y, equal_nan=True) for x, y in zip(base, tight))


def test_one_list_over_two_devices_in_process():
    """sxs_compute_saxs_scores with SXS_CUDA_DEVICES=0,1 (one host thread per device, each extracting the rows of its own
    z range — the replacement of `mpirun -np 2 correlate`, tools/correlate.c:140-147,295-356) returns the bits one
    device returns, rows outside the z table untouched; needs two GPUs (gpurun --gpus 2)"""
    if capi.device_count() < 2:
        pytest.skip("one CUDA device only")
    G = np.load(os.path.join(GOLD, "golden_4g9s.npz"))
    R = np.load(os.path.join(GOLD, "golden_real70k.npz"))
    q, L = G["qvals"], int(G["L"])
    nb, N = L + 1, 2 * L + 1
    idx = R["index"].copy()
    idx = np.concatenate([idx[:40000], np.array([-3], dtype=idx.dtype), idx[40000:]])
    init = (np.full(len(idx), 7.0), np.full(len(idx), 8.0), np.full(len(idx), 9.0))
    old = os.environ.get("SXS_CUDA_DEVICES")
    try:
        os.environ["SXS_CUDA_DEVICES"] = "0"
        one = capi.scores(idx, G["rec_coef"], G["lig_coef"], G["a"], G["scal"], q, R["zvals"], L, init=init)
        os.environ["SXS_CUDA_DEVICES"] = "0,1"
        two = capi.scores(idx, G["rec_coef"], G["lig_coef"], G["a"], G["scal"], q, R["zvals"], L, init=init)
        both64 = capi.scores(idx.astype(np.int64), G["rec_coef"], G["lig_coef"], G["a"], G["scal"], q, R["zvals"], L, init=init)
    finally:
        if old is None:
            os.environ.pop("SXS_CUDA_DEVICES", None)
        else:
            os.environ["SXS_CUDA_DEVICES"] = old
    assert all(np.array_equal(x, y) for x, y in zip(one, two))
    assert all(np.array_equal(x, y) for x, y in zip(one, both64))
    assert one[0][40000] == 7.0 and two[1][40000] == 8.0 and two[2][40000] == 9.0
    keep = np.arange(len(idx)) != 40000
    parity.check("real list over two devices", tuple(x[keep] for x in two), (R["scores"], R["c1"], R["c2"]),
                 sens=(R["sens_scores"], R["sens_c1"], R["sens_c2"]))


def test_translation_kernels_agree_bit_for_bit():
    """K2c: the register-tiled translation (4 x 2 outputs per thread) writes the same St as the one-output-per-thread
    form (SXS_TRANSLATE_V1) — scores, c1, c2 are equal to the last bit — at L = 15 and at L = 30 / Q = 100"""
    def both(plan, idx):
        os.environ.pop("SXS_TRANSLATE_V1", None)
        new = plan.score(idx)
        os.environ["SXS_TRANSLATE_V1"] = "1"
        try:
            old = plan.score(idx)
        finally:
            os.environ.pop("SXS_TRANSLATE_V1", None)
        return new, old

    G = np.load(os.path.join(GOLD, "golden_4g9s.npz"))
    R = np.load(os.path.join(GOLD, "golden_real70k.npz"))
    plan = capi.Plan(int(G["L"]), G["qvals"])
    plan.set_molecules(G["rec_coef"], G["lig_coef"])
    plan.set_experiment(G["a"], G["scal"][1], G["scal"][2])
    plan.set_translations(R["zvals"])
    new, old = both(plan, R["index"][:20000])
    plan.close()
    assert all(np.array_equal(x, y, equal_nan=True) for x, y in zip(new, old))
    assert np.max(np.abs(new[0] / R["scores"][:20000] - 1)) < TOL

    F = np.load(os.path.join(GOLD, "golden_l30.npz"))
    L, q = int(F["L"]), F["qvals"]
    A, _, _ = capi.expand(wl.MAP_PATH, F["rec_xyz"], names(F["rec_res"]), names(F["rec_atm"]), F["rec_radius"], q, L,
                          sa=F["rec_sa"], water_mode=1)
    B, _, _ = capi.expand(wl.MAP_PATH, F["lig_xyz"], names(F["lig_res"]), names(F["lig_atm"]), F["lig_radius"], q, L,
                          sa=F["lig_sa"], water_mode=1)
    plan = capi.Plan(L, q)
    plan.set_molecules(A, B)
    plan.set_experiment(F["a"], F["scal"][1], F["scal"][2])
    plan.set_translations(F["zvals"])
    new, old = both(plan, F["index"])
    plan.close()
    assert all(np.array_equal(x, y, equal_nan=True) for x, y in zip(new, old))
    assert np.max(np.abs(new[0] / F["scores"] - 1)) < TOL


def test_grouped_cross_kernel_is_bit_identical():
    """K3 threads that take 2 or 4 points differing only in a2 (same cell, g1, g2: shared operands and inner sums) write
    the same cross terms as one point per thread: scores, c1, c2 equal bit for bit on a dense list, on a sparse one and
    across chunk boundaries; unset, the run length decides"""
    def run(plan, idx, group):
        old = os.environ.get("SXS_CROSS_GROUP")
        if group is None:
            os.environ.pop("SXS_CROSS_GROUP", None)
        else:
            os.environ["SXS_CROSS_GROUP"] = str(group)
        try:
            return plan.score(idx), plan.stats()
        finally:
            if old is None:
                os.environ.pop("SXS_CROSS_GROUP", None)
            else:
                os.environ["SXS_CROSS_GROUP"] = old

    w = wl.make("cfg3_3k+1.5k_L15_Q50_70kx64z", nrot=20000, nz=3)
    q, L = w["qvals"], w["L"]
    nb, N = L + 1, 2 * L + 1
    A, _, _ = capi.expand(wl.MAP_PATH, w["rec"]["xyz"], w["rec"]["res"], w["rec"]["atm"], w["rec"]["radius"], q, L,
                          sa=w["rec"]["sa"], water_mode=1)
    B, _, _ = capi.expand(wl.MAP_PATH, w["lig"]["xyz"], w["lig"]["res"], w["lig"]["atm"], w["lig"]["radius"], q, L,
                          sa=w["lig"]["sa"], water_mode=1)
    eq, ei, ee = wl.experimental_curve(A, B, q)
    a, scal = capi.opt_params(eq, ei, ee, q, wl.mean_radius(w["rec"], w["lig"]))
    idx = w["index"].astype(np.int64)
    # fold the (beta1, beta2) digits onto 3 x 2 values: cells of ~3300 rows, ~3.5 rows per (g1, g2) pair
    n3 = N ** 3
    ang, cell = idx % n3, idx // n3
    b2, b1, z = cell % nb, (cell // nb) % nb, cell // (nb * nb)
    dense = (((z * nb + (b1 % 3) + 5) * nb + (b2 % 2) + 7) * n3 + ang).astype(np.int32)
    plan = capi.Plan(L, q)
    plan.set_molecules(A, B)
    plan.set_experiment(a, scal[1], scal[2])
    plan.set_translations(w["zvals"])
    for lst in (dense, w["index"]):
        base, st1 = run(plan, lst, 1)
        assert st1["cross_groups"] == 0
        for k in (2, 4):
            got, stk = run(plan, lst, k)
            assert 0 < stk["cross_groups"] <= stk["points"]
            assert all(np.array_equal(x, y, equal_nan=True) for x, y in zip(base, got)), "group=%d" % k
    auto, sta = run(plan, dense, None)
    assert 0 < sta["cross_groups"] < 0.5 * sta["points"]          # ~3.5 per pair -> K = 4
    assert all(np.array_equal(x, y, equal_nan=True) for x, y in zip(run(plan, dense, 1)[0], auto))
    _, sts = run(plan, w["index"][::7], None)                       # thinned list: nearly every pair once -> plain kernel
    assert sts["cross_groups"] == 0
    plan.close()
    old = os.environ.get("SXS_CUDA_X_GB")
    os.environ["SXS_CUDA_X_GB"] = "0.002"                           # ~1000 points per chunk: chunk boundaries cut runs
    try:
        small = capi.Plan(L, q)
        small.set_molecules(A, B)
        small.set_experiment(a, scal[1], scal[2])
        small.set_translations(w["zvals"])
        sub = dense[:30000]
        b1_, _ = run(small, sub, 1)
        for k in (2, 4):
            gk, _ = run(small, sub, k)
            assert all(np.array_equal(x, y, equal_nan=True) for x, y in zip(b1_, gk))
        small.close()
    finally:
        if old is None:
            os.environ.pop("SXS_CUDA_X_GB", None)
        else:
            os.environ["SXS_CUDA_X_GB"] = old


def test_dense_scan_topk():
    """every grid point of one z step (16 x 16 cells x 31^3 = 7.6 M points) is scored on the device and
    the best 64 come back.  The scan runs the DENSE form of K3 (k_cross_dense: the inner sums of a (g1, g2) pair serve all
    31 values of a2); it must name the points the list form names (SXS_SCAN_LIST=1: the same scan through the list
    kernels), carry what the list API gives for those indices to rounding, and its K3 must be >= 1.5x faster (measured
    2.2x: 41 against 91 ms per z with the operands of every m staged in shared memory; 47 ms with plain loads)."""
    G = np.load(os.path.join(GOLD, "golden_4g9s.npz"))
    q, L = G["qvals"], int(G["L"])
    nb, N = L + 1, 2 * L + 1
    plan = capi.Plan(L, q)
    plan.set_molecules(G["rec_coef"], G["lig_coef"])
    plan.set_experiment(G["a"], G["scal"][1], G["scal"][2])
    plan.set_translations(np.array([38.0, 40.0]))
    k = 64
    plan.set_profiling(True)
    idx, s, c1, c2 = plan.scan_topk(k, z_lo=1, z_hi=2)
    t_dense = plan.kernel_times()
    per_z = nb * nb * N ** 3
    assert plan.stats()["points"] == per_z
    assert (idx >= per_z).all() and (idx < 2 * per_z).all() and len(np.unique(idx)) == k
    assert np.all(np.diff(s) >= 0) and np.isfinite(s).all()
    os.environ["SXS_SCAN_LIST"] = "1"
    try:
        idx_l, s_l, c1_l, c2_l = plan.scan_topk(k, z_lo=1, z_hi=2)
        t_list = plan.kernel_times()
    finally:
        os.environ.pop("SXS_SCAN_LIST", None)
    plan.set_profiling(False)
    print("dense scan of one z: K3 %.1f ms dense form, %.1f ms list form; K4 %.1f / %.1f ms"
          % (t_dense["cross"][0], t_list["cross"][0], t_dense["fit"][0], t_list["fit"][0]))
    assert t_dense["cross"][0] * 1.5 <= t_list["cross"][0]
    assert np.array_equal(np.sort(idx), np.sort(idx_l))
    o, ol = np.argsort(idx), np.argsort(idx_l)
    parity.check("dense form vs list form, top-%d" % k, (s[o], c1[o], c2[o]), (s_l[ol], c1_l[ol], c2_l[ol]))
    ls, lc1, lc2 = plan.score(idx)
    parity.check("dense scan vs list API on the same indices", (s, c1, c2), (ls, lc1, lc2))
    sample = np.random.default_rng(5).integers(per_z, 2 * per_z, 200000)
    ss, _, _ = plan.score(sample)
    better = sample[ss < s[-1] * (1 - 1e-9)]
    assert np.isin(better, idx).all()
    assert ss.min() >= s[0] * (1 - 1e-9)
    plan.close()


@pytest.mark.skipif(not refso.available(), reason="compiled reference (oracle/_ref) not present")
def test_dense_scan_topk_against_the_reference():
    """the dense scan against the ORACLE: (i) the k best points it reports carry the reference's (chi, c1, c2);
    (ii) top-ness inside one cell: the reference scores ALL 29 791 grid points of the cell that holds the best point
    (its skip = 0 mode computes exactly this grid, src/fftsaxs.c:867-872) — the scan's members from that cell are the
    reference's points below the k-th score, and no other point of the cell is"""
    G = np.load(os.path.join(GOLD, "golden_4g9s.npz"))
    q, L = G["qvals"], int(G["L"])
    nb, N = L + 1, 2 * L + 1
    zv = np.array([40.0])
    plan = capi.Plan(L, q)
    plan.set_molecules(G["rec_coef"], G["lig_coef"])
    plan.set_experiment(G["a"], G["scal"][1], G["scal"][2])
    plan.set_translations(zv)
    k = 48
    idx, s, c1, c2 = plan.scan_topk(k, z_lo=0, z_hi=1)
    plan.close()
    want = refso.scores(idx.astype(np.int32), G["rec_coef"], G["lig_coef"], G["a"], G["scal"], q, zv, L)
    parity.check("top-%d of a dense scan, reference on the same points" % k, (s, c1, c2), want)
    cell = int(idx[0]) // N ** 3
    allpts = (cell * N ** 3 + np.arange(N ** 3)).astype(np.int32)
    rs, _, _ = refso.scores(allpts, G["rec_coef"], G["lig_coef"], G["a"], G["scal"], q, zv, L)
    mine_in_cell = np.sort(idx[idx // N ** 3 == cell])
    thr = s[-1]
    ref_in_cell = np.sort(allpts[rs < thr * (1 - 1e-9)].astype(np.int64))
    assert np.isin(ref_in_cell, mine_in_cell).all()                       # nothing better was missed
    assert (rs[np.isin(allpts, mine_in_cell)] <= thr * (1 + 1e-9)).all()    # everything reported belongs
    assert rs.min() >= s[0] * (1 - 1e-9)


def test_ft_rows_to_indices_on_the_device_equal_the_host_route():
    """SURVEY 8f-2 on the device: k_ft_rows_to_index (one thread per ft row: sxs_ft2euler, the three-decimal text round
    trip as exact arithmetic, z lookup, snapping) gives, row for row, the flat indices of the host route
    sxs_ft_rows_to_indices64 (itself pinned on the reference's Euler file byte for byte, tests/test_cpu_host.py) —
    2 M random rows incl. rows off the z table, bad rotation ids, and translations on the axes (degenerate angles)."""
    import time
    rng = np.random.default_rng(21)
    L, nrot, nrows = 15, 70000, 2_000_000
    qn = rng.normal(size=(nrot, 4))
    qn /= np.linalg.norm(qn, axis=1)[:, None]
    w, x, y, z = qn.T
    R = np.stack([1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w), 2 * (x * y + z * w),
                  1 - 2 * (x * x + z * z), 2 * (y * z - x * w), 2 * (x * z - y * w), 2 * (y * z + x * w),
                  1 - 2 * (x * x + y * y)], 1)
    rot_id = rng.integers(0, nrot, nrows).astype(np.int32)
    u = rng.normal(size=(nrows, 3))
    u /= np.linalg.norm(u, axis=1)[:, None]
    ref_lig = np.array([3.0, -2.0, 1.5])
    dist = rng.uniform(-2.0, 90.0, nrows)
    trans = u * dist[:, None] - ref_lig
    # translations whose rounded z lands exactly on table values, and on the coordinate axes
    trans[:1000] = np.round(trans[:1000] + ref_lig) - ref_lig
    trans[1000:1003] = np.array([[0, 0, 30.0], [0, 0, -30.0], [30.0, 0, 0]]) - ref_lig
    zvals = np.arange(1.0, 80.001, 1.0)
    t0 = time.time()
    index, ft_id, order = capi.ft_rows_to_indices(rot_id, trans, R, ref_lig, zvals, L)
    t_host = time.time() - t0
    want = np.full(nrows, -1, dtype=np.int64)
    want[order] = index
    capi.cuda_ft_rows_to_indices(rot_id[:10], trans[:10], R, ref_lig, zvals, L)   # context + module load
    t0 = time.time()
    got = capi.cuda_ft_rows_to_indices(rot_id, trans, R, ref_lig, zvals, L)
    t_dev = time.time() - t0
    # degenerate rows (0 / 0 angles: translation on the z axis) carry the tool's (int)round(nan) digits on the host — a
    # negative index that the scoring layer leaves untouched — and -1 on the device
    diff = np.flatnonzero((got != want) & ((want >= 0) | (got >= 0)))
    assert got[1000] == -1 and want[1000] < 0            # (0, 0, +30): b1 = 0, g1 = acos(0 / 0)
    print("ft rows -> indices, %d rows: host threads %.3f s (%.2f M rows/s), device incl. copies %.3f s (%.2f M rows/s); "
          "%d kept, %d rows differ" % (nrows, t_host, nrows / t_host / 1e6, t_dev, nrows / t_dev / 1e6, (want >= 0).sum(), len(diff)))
    assert 0 < (want >= 0).sum() < nrows
    assert len(diff) == 0, (diff[:10], got[diff[:10]], want[diff[:10]])
    bad = rot_id[:5].copy()
    bad[2] = nrot
    assert capi.cuda_ft_rows_to_indices(bad, trans[:5], R, ref_lig, zvals, L)[2] == -2
    # L = 30: 64-bit indices
    i30, _, o30 = capi.ft_rows_to_indices(rot_id[:200000], trans[:200000], R, ref_lig, zvals, 30)
    w30 = np.full(200000, -1, dtype=np.int64)
    w30[o30] = i30
    g30 = capi.cuda_ft_rows_to_indices(rot_id[:200000], trans[:200000], R, ref_lig, zvals, 30)
    ok30 = (w30 >= 0) | (g30 >= 0)
    assert np.array_equal(g30[ok30], w30[ok30]) and w30.max() > 2 ** 31


@pytest.mark.parametrize("L,Q", [(3, 7), (6, 6), (9, 7), (17, 5)])
def test_dense_scan_equals_the_list_path_at_other_orders(L, Q):
    """the dense form of K3 is instantiated per range of L (lanes per pair, a2 pairs per lane, block size): at every one
    of them the k best points of a dense scan carry the scores the LIST kernels give for the same indices, and 32- and
    64-bit index lists give identical results (synthetic molecules: 120 + 60 atoms of the bench workload)"""
    base = wl.make("cfg3_3k+1.5k_L15_Q50_70kx64z", nrot=1, nz=1)
    rec, lig = base["rec"], base["lig"]
    q = capi.mkarray(0.0, 0.4, Q)
    nb, N = L + 1, 2 * L + 1
    r, l = slice(0, 120), slice(0, 60)
    A, _, _ = capi.expand(wl.MAP_PATH, rec["xyz"][r], rec["res"][r], rec["atm"][r], rec["radius"][r], q, L, sa=rec["sa"][r], water_mode=1)
    B, _, _ = capi.expand(wl.MAP_PATH, lig["xyz"][l], lig["res"][l], lig["atm"][l], lig["radius"][l], q, L, sa=lig["sa"][l], water_mode=1)
    eq = np.linspace(0.0, 0.45, 40)
    ei = 1e4 * np.exp(-eq * eq * 60.0) + 50.0
    a, scal = capi.opt_params(eq, ei, 0.03 * ei, q, 1.6)
    plan = capi.Plan(L, q)
    plan.set_molecules(A, B)
    plan.set_translations(np.array([20.0, 21.0, 22.0]))
    plan.set_experiment(a, scal[1], scal[2])
    per_z = nb * nb * N ** 3
    rng = np.random.default_rng(L)
    idx = rng.integers(0, 3 * per_z, 400)
    idx[::50] = -1                                   # rows that stay untouched
    idx[1::50] = 5 * per_z                           # z digit off the table
    s32 = plan.score(idx.astype(np.int32))
    s64 = plan.score(idx.astype(np.int64))
    assert all(np.array_equal(x, y) for x, y in zip(s32, s64)) and np.isfinite(s32[0]).all()
    assert (s32[0][::50] == 0).all() and (s32[0][1::50] == 0).all() and (s32[0][2::50] > 0).all()
    k = 16
    ti, ts, tc1, tc2 = plan.scan_topk(k, z_lo=1, z_hi=2)
    assert (ti >= per_z).all() and (ti < 2 * per_z).all() and len(np.unique(ti)) == k and np.all(np.diff(ts) >= 0)
    ls, lc1, lc2 = plan.score(ti)
    parity.check("dense scan vs list kernels, L = %d" % L, (ts, tc1, tc2), (ls, lc1, lc2))
    sample = rng.integers(per_z, 2 * per_z, 20000)
    ss, _, _ = plan.score(sample)
    assert ss.min() >= ts[0] * (1 - 1e-9) and np.isin(sample[ss < ts[-1] * (1 - 1e-9)], ti).all()
    plan.close()


@pytest.mark.parametrize("L", [10, 20, 40])
def test_config5_orders_against_reference(L):
    """BASELINE config 5 sweeps lmax over 10, 15, 20, 30, 40: besides 15 (the goldens) and 30 (golden_l30) the orders 10,
    20 and 40 against the compiled reference — 35 to 49 poses in two or three cells, one of them with >= 30 rows (the
    reference's FFT branch), 32- and 64-bit lists; fixture by tests/golden/make_golden_cfg5.py with the reference's own
    FMA-build answers as noise floor"""
    F = np.load(os.path.join(GOLD, "golden_cfg5_orders.npz"))
    g = lambda k: F["L%d_%s" % (L, k)]  # noqa: E731
    q = g("qvals")
    rec = wl.make_molecule(300, 100 + L)
    lig = wl.make_molecule(150, 200 + L)
    rec["xyz"] -= 0.5 * (rec["xyz"].min(0) + rec["xyz"].max(0))
    lig["xyz"] -= lig["xyz"].mean(0)
    A, _, _ = capi.expand(wl.MAP_PATH, rec["xyz"], rec["res"], rec["atm"], rec["radius"], q, L, sa=rec["sa"], water_mode=1)
    B, _, _ = capi.expand(wl.MAP_PATH, lig["xyz"], lig["res"], lig["atm"], lig["radius"], q, L, sa=lig["sa"], water_mode=1)
    for mine, ref in ((A, g("coefA_sample")), (B, g("coefB_sample"))):
        assert np.max(np.abs(mine[:, ::7, ::53] - ref)) / np.abs(ref).max() < 1e-9
    want = (g("scores"), g("c1"), g("c2"))
    sens = (g("sens_scores"), g("sens_c1"), g("sens_c2"))
    got = capi.scores(g("index"), A, B, g("a"), g("scal"), q, g("zvals"), L)
    parity.check("config 5 order L = %d, %d poses" % (L, len(want[0])), got, want, sens=sens)
    got64 = capi.scores(g("index").astype(np.int64), A, B, g("a"), g("scal"), q, g("zvals"), L)
    assert all(np.array_equal(x, y) for x, y in zip(got, got64))
