"""Fixtures for the other orders of BASELINE config 5 (L = 10, 20, 40): small synthetic molecules, a few dozen poses in a
few cells (one of them with >= 30 rows, the reference's FFT branch), scored by the compiled reference and by its
FMA-contracted build (noise floor, tests/parity.py).  Run in the build container:
    python tests/golden/make_golden_cfg5.py
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import refso  # noqa: E402
from libfmftsaxs_b200 import workload as wl  # noqa: E402

CASES = [(10, 50, [(0, 3, 7, 36), (1, 9, 2, 8), (1, 0, 10, 5)]),
         (20, 50, [(0, 5, 14, 33), (1, 20, 1, 6)]),
         (40, 24, [(0, 11, 29, 31), (1, 33, 40, 4)])]


def main():
    out = {}
    for L, Q, cells in CASES:
        rec = wl.make_molecule(300, 100 + L)
        lig = wl.make_molecule(150, 200 + L)
        rec["xyz"] -= 0.5 * (rec["xyz"].min(0) + rec["xyz"].max(0))
        lig["xyz"] -= lig["xyz"].mean(0)
        q = wl.make_qvals(Q)
        A, _, _ = refso.expand(wl.MAP_PATH, rec["xyz"], rec["res"], rec["atm"], rec["radius"], q, L, sa=rec["sa"], water_mode=1)
        B, _, _ = refso.expand(wl.MAP_PATH, lig["xyz"], lig["res"], lig["atm"], lig["radius"], q, L, sa=lig["sa"], water_mode=1)
        eq, ei, ee = wl.experimental_curve(A, B, q)
        a, scal = refso.opt_params(eq, ei, ee, q, wl.mean_radius(rec, lig))
        zv = np.array([22.0, 31.0])
        nb, N = L + 1, 2 * L + 1
        rng = np.random.default_rng(L)
        idx = []
        for (z, b1, b2, rows) in cells:
            for _ in range(rows):
                a2, g1, g2 = rng.integers(0, N, 3)
                idx.append((((((z * nb + b1) * nb + b2) * N + a2) * N + g1) * N + g2))
        idx = np.array(idx, dtype=np.int64)
        assert idx.max() < 2 ** 31
        idx = idx.astype(np.int32)
        t = time.time()
        s, c1, c2 = refso.scores(idx, A, B, a, scal, q, zv, L)
        t1 = time.time() - t
        ss, sc1, sc2 = refso.scores(idx, A, B, a, scal, q, zv, L, so=refso.SENS_SO)
        print("L = %d, Q = %d: %d poses in %d cells, reference %.1f s; chi %.3f..%.3f, c2 %.3f..%.3f" %
              (L, Q, len(idx), len(cells), t1, s.min(), s.max(), c2.min(), c2.max()), flush=True)
        tag = "L%d_" % L
        # the coefficient tables themselves are not stored (MBs at L = 40): the test expands the same seeded molecules on
        # the device (K1 is held to 1e-9 of scale elsewhere) and checks this sample
        for k, v in dict(qvals=q, zvals=zv, index=idx, a=a, scal=scal, coefA_sample=A[:, ::7, ::53], coefB_sample=B[:, ::7, ::53],
                         scores=s, c1=c1, c2=c2, sens_scores=ss, sens_c1=sc1, sens_c2=sc2).items():
            out[tag + k] = v
    np.savez_compressed(os.path.join(HERE, "golden_cfg5_orders.npz"), orders=np.array([c[0] for c in CASES]), **out)
    print("%.1f KB" % (os.path.getsize(os.path.join(HERE, "golden_cfg5_orders.npz")) / 1e3))


if __name__ == "__main__":
    main()
