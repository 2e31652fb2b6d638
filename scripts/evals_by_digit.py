#!/usr/bin/env python
"""How the number of objective evaluations per fit depends on where a point lies on the grid (config 3 workload):
cross terms of a sample of poses per z step -> K4 alone -> evaluations.  Looks for a cheap predictor of long fits
(the tail of a K4 launch is the longest fit that starts late)."""
import os, sys
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import bench
from libfmftsaxs_b200 import capi

def main():
    def native_cross(idx, coefA, coefB, q, zv, L):
        pl = capi.Plan(L, q); pl.set_molecules(coefA, coefB); pl.set_experiment(np.ones(6 * len(q)), 1.0, 1.0); pl.set_translations(zv)
        X = pl.cross_terms(idx); pl.close(); return X
    w = bench.build_inputs(capi.expand, capi.opt_params, native_cross, 0, 2000, None)
    L, q, zvals, idx = w["L"], w["qvals"], w["zvals"], w["index"]
    nb, N = L + 1, 2 * L + 1
    plan = capi.Plan(L, q); plan.set_molecules(w["coefA"], w["coefB"]); plan.set_experiment(w["a"], w["scal"][1], w["scal"][2]); plan.set_translations(zvals)
    rng = np.random.default_rng(1)
    sel = rng.choice(len(idx), 60000, replace=False)
    ii = idx[sel].astype(np.int64)
    X = plan.cross_terms(ii.astype(np.int32))
    out = capi.cuda_fit_profiles(X, w["a"], q, w["scal"][1], w["scal"][2])
    ev = out[:, 3]
    zd = ii // (nb * nb * N ** 3)
    b1 = (ii // (nb * N ** 3)) % nb
    b2 = (ii // (N ** 3)) % nb
    print("all: mean %.2f p50 %d p90 %d p99 %d max %d" % (ev.mean(), np.percentile(ev, 50), np.percentile(ev, 90), np.percentile(ev, 99), ev.max()))
    for name, d, nd in (("z", zd, len(zvals)), ("b1", b1, nb), ("b2", b2, nb)):
        m = np.array([ev[d == k].mean() if (d == k).any() else np.nan for k in range(nd)])
        p = np.array([(ev[d == k] >= 35).mean() if (d == k).any() else np.nan for k in range(nd)])
        print(name, "mean evals per digit:", np.round(m, 1)[:: max(1, nd // 16)])
        print(name, "share of fits with >= 35 evals:", np.round(p, 3)[:: max(1, nd // 16)])
    # chi, c1, c2 of the result and the peak-scaled I(0)
    for name, v in (("chi", out[:, 0]), ("c2", out[:, 2]), ("c1", out[:, 1]), ("X00", X[:, 0, 0])):
        print("corr(evals, %s) = %.3f" % (name, np.corrcoef(ev, v)[0, 1]))
    plan.close()

if __name__ == "__main__":
    main()
