"""Adds the answers of the reference's SENSITIVITY build (oracle/_ref/libsxsref_fma.so: same unmodified sources, FMA
contraction on) to the real-data fixtures: golden_4g9s.npz (z6_sens_*) and golden_real70k.npz (sens_*).  They are the
reference's own reaction to a 1e-16 perturbation of its arithmetic, printed next to every parity result.
Run in the build container:  python tests/golden/add_sensitivity.py
"""
import os
import sys
from multiprocessing import Pool

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import refso  # noqa: E402

L = 15


def _one(job):
    idx, A, B, a, scal, q, zv = job
    return refso.scores(idx, A, B, a, scal, q, zv, L, so=refso.SENS_SO)


def sens(idx, zvals, G):
    nb, N = L + 1, 2 * L + 1
    zdig = idx // (nb * nb * N ** 3)
    zs = np.unique(zdig)
    jobs = [(idx[zdig == z], G["rec_coef"], G["lig_coef"], G["a"], G["scal"], G["qvals"], zvals) for z in zs]
    with Pool(8) as pool:
        res = pool.map(_one, jobs)
    out = [np.zeros(len(idx)) for _ in range(3)]
    for z, r in zip(zs, res):
        for k in range(3):
            out[k][zdig == z] = r[k]
    return out


def main():
    p = os.path.join(HERE, "golden_4g9s.npz")
    G = dict(np.load(p))
    s = sens(G["z6_index"], G["z6_zvals"], G)
    G.update(z6_sens_scores=s[0], z6_sens_c1=s[1], z6_sens_c2=s[2])
    np.savez_compressed(p, **G)
    p = os.path.join(HERE, "golden_real70k.npz")
    R = dict(np.load(p))
    s = sens(R["index"], R["zvals"], G)
    R.update(sens_scores=s[0], sens_c1=s[1], sens_c2=s[2])
    np.savez_compressed(p, **R)
    d2 = np.abs(s[2] - R["c2"]) / np.maximum(np.abs(R["c2"]), 1e-3)
    print("70k rows: reference vs its FMA build, rows beyond 1e-6 in c2: %d, max %.2e" % ((d2 > 1e-6).sum(), d2.max()))


if __name__ == "__main__":
    main()
