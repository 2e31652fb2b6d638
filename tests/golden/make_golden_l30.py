"""Fixture for BASELINE config 4's shape (L = 30, 100 q points, 64-bit-safe indices): small synthetic molecules,
a few poses in two cells, scored by the compiled reference.  Run in the build container:
    python tests/golden/make_golden_l30.py
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

L, Q = 30, 100
rec = wl.make_molecule(300, 11)
lig = wl.make_molecule(150, 12)
rec["xyz"] -= 0.5 * (rec["xyz"].min(0) + rec["xyz"].max(0))
lig["xyz"] -= lig["xyz"].mean(0)
q = wl.make_qvals(Q)
A, rmA, _ = refso.expand(wl.MAP_PATH, rec["xyz"], rec["res"], rec["atm"], rec["radius"], q, L, sa=rec["sa"], water_mode=1)
B, rmB, _ = refso.expand(wl.MAP_PATH, lig["xyz"], lig["res"], lig["atm"], lig["radius"], q, L, sa=lig["sa"], water_mode=1)
eq, ei, ee = wl.experimental_curve(A, B, q)
a, scal = refso.opt_params(eq, ei, ee, q, wl.mean_radius(rec, lig))
zv = np.array([20.0, 35.5])
nb, N = L + 1, 2 * L + 1
rng = np.random.default_rng(5)
dig = [(0, 7, 22), (1, 15, 3)]
idx = []
for (z, b1, b2) in dig:
    for _ in range(3):
        a2, g1, g2 = rng.integers(0, N, 3)
        idx.append((((((z * nb + b1) * nb + b2) * N + a2) * N + g1) * N + g2))
idx = np.array(idx, dtype=np.int32)  # 2 z steps fit 32 bits at L = 30 (limit is 9)
t = time.time()
s, c1, c2 = refso.scores(idx, A, B, a, scal, q, zv, L)
print("reference: %.1fs" % (time.time() - t), s, c1, c2)
np.savez_compressed(os.path.join(HERE, "golden_l30.npz"), L=L, qvals=q, zvals=zv, index=idx, a=a, scal=scal,
                    rec_xyz=rec["xyz"], rec_radius=rec["radius"], rec_sa=rec["sa"], rec_res=np.array(rec["res"], dtype="S8"),
                    rec_atm=np.array(rec["atm"], dtype="S8"), lig_xyz=lig["xyz"], lig_radius=lig["radius"], lig_sa=lig["sa"],
                    lig_res=np.array(lig["res"], dtype="S8"), lig_atm=np.array(lig["atm"], dtype="S8"),
                    scores=s, c1=c1, c2=c2, coefA_sample=A[:, ::25, ::97], coefB_sample=B[:, ::25, ::97])
