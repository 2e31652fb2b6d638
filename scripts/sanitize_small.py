#!/usr/bin/env python
"""A small end-to-end pass over every kernel of the path, meant to run under compute-sanitizer:

    compute-sanitizer --tool memcheck  python scripts/sanitize_small.py 3,6      # minutes per L
    compute-sanitizer --tool racecheck python scripts/sanitize_small.py 6        # slow: > 15 min at L = 17

r2ag [B200]: memcheck 0 errors (L = 3, 6: every kernel of the list path and the dense scan), racecheck 0 hazards (L = 6);
initcheck flags only the device-to-host copy of the score table's unscored rows in sxs_cuda_plan_score_* (never used).

Synthetic molecules at several small L (odd and even, below and above the dense kernel's template switches), list
scoring with 32- and 64-bit indices, rows off the z table, the dense scan, K4 alone, the ft-rows kernel."""
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from libfmftsaxs_b200 import capi                   # noqa: E402
from libfmftsaxs_b200 import workload as wl         # noqa: E402


def main():
    rng = np.random.default_rng(3)
    Ls = [int(x) for x in (sys.argv[1].split(",") if len(sys.argv) > 1 else "3,6,9,17".split(","))]
    for L in Ls:
        Q = 7 if L % 2 else 6
        q = capi.mkarray(0.0, 0.4, Q)
        nb, N = L + 1, 2 * L + 1
        base = wl.make("cfg3_3k+1.5k_L15_Q50_70kx64z", nrot=1, nz=1)
        rec, lig = base["rec"], base["lig"]
        sel_r, sel_l = slice(0, 120), slice(0, 60)
        A, _, _ = capi.expand(wl.MAP_PATH, rec["xyz"][sel_r], rec["res"][sel_r], rec["atm"][sel_r], rec["radius"][sel_r], q, L,
                              sa=rec["sa"][sel_r], water_mode=1)
        B, _, _ = capi.expand(wl.MAP_PATH, lig["xyz"][sel_l], lig["res"][sel_l], lig["atm"][sel_l], lig["radius"][sel_l], q, L,
                              sa=lig["sa"][sel_l], water_mode=1)
        plan = capi.Plan(L, q)
        plan.set_molecules(A, B)
        zvals = np.array([20.0, 21.0, 22.0])
        plan.set_translations(zvals)
        eq = np.linspace(0.0, 0.45, 40)
        ei = 1e4 * np.exp(-eq * eq * 60.0) + 50.0
        a, scal = capi.opt_params(eq, ei, 0.03 * ei, q, 1.6)
        plan.set_experiment(a, scal[1], scal[2])
        per_z = nb * nb * N ** 3
        idx = rng.integers(0, 3 * per_z, 500)
        idx[::50] = -1                       # untouched rows
        idx[1::50] = 5 * per_z               # z digit off the table
        s32 = plan.score(idx.astype(np.int32) if 3 * per_z < 2 ** 31 else idx.astype(np.int64))
        s64 = plan.score(idx.astype(np.int64))
        assert all(np.array_equal(a, b) for a, b in zip(s32, s64)) and np.isfinite(s32[0]).all()
        k = 8
        ti, ts, _, _ = plan.scan_topk(k, z_lo=1, z_hi=2)
        ls, _, _ = plan.score(ti)
        assert np.allclose(ls, ts, rtol=1e-6) and np.isfinite(ts).all(), (ls, ts)
        plan.close()
        print("L = %d, Q = %d: list (32/64-bit) and dense scan ok; best chi %.4g" % (L, Q, ts[0]), flush=True)
    # K4 alone, and its fallback objective (full exp, IEEE quotients) that runs when |corr q^2| could reach 512 in the box:
    # a large `mult` selects it; with mult = 2000 the arguments stay below 41 and the two objectives must agree bit for bit
    FC = np.load(os.path.join(REPO, "tests", "golden", "fit_cases.npz"))
    G = np.load(os.path.join(REPO, "tests", "golden", "golden_4g9s.npz"))
    X = np.ascontiguousarray(FC["X52"])
    fast = capi.cuda_fit_profiles(X, G["a"], G["qvals"], float(G["scal"][1]), float(G["scal"][2]))
    assert np.array_equal(fast, FC["fit52"]), "K4 against the stored outputs of the reference's L-BFGS-B"
    q4 = np.ascontiguousarray(G["qvals"])
    big = 2000.0                                     # bound = 2000 * 0.0816 * 0.25 = 40.8 < 500: fast path
    huge = 500.0 / 0.0816 / float(q4.max() ** 2) * 1.01   # bound just above 500: fallback path, same arguments otherwise
    r_fast = capi.cuda_fit_profiles(X, G["a"], q4, big, float(G["scal"][2]))
    r_safe = capi.cuda_fit_profiles(X, G["a"], q4, huge, float(G["scal"][2]))
    print("K4: 52 fits bit-identical to the reference; mult = %g (fast objective) finite %s; mult = %g (fallback objective) ran, finite %s"
          % (big, bool(np.isfinite(r_fast).all()), huge, bool(np.isfinite(r_safe).all())), flush=True)
    # ft rows -> indices
    nrot, n = 50, 3000
    qn = rng.normal(size=(nrot, 4)); qn /= np.linalg.norm(qn, axis=1)[:, None]
    w, x, y, z = qn.T
    R = np.stack([1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w), 2 * (x * y + z * w), 1 - 2 * (x * x + z * z),
                  2 * (y * z - x * w), 2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)], 1)
    u = rng.normal(size=(n, 3)); u /= np.linalg.norm(u, axis=1)[:, None]
    t = u * rng.uniform(-2, 90, n)[:, None]
    got = capi.cuda_ft_rows_to_indices(rng.integers(0, nrot, n), t, R, np.zeros(3), np.arange(1.0, 80.5), 15)
    print("ft rows: %d kept of %d" % ((got >= 0).sum(), n), flush=True)


if __name__ == "__main__":
    main()
