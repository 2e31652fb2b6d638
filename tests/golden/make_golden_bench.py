"""Fixtures on the BENCHMARKED workloads (BASELINE configs 3 and 4 exactly as bench.py generates them), scored by the
compiled reference.  Run in the build container only (needs oracle/_ref/libsxsref.so, i.e. /root/reference):

    python tests/golden/make_golden_bench.py cfg3        -> golden_cfg3_slabs.npz   (~1 min)
    python tests/golden/make_golden_bench.py cfg4        -> golden_cfg4_cells.npz   (~2 min, 8 processes)

cfg3: the full molecules (3000 + 1500 atoms), L = 15, Q = 50, the experimental curve of bench.build_inputs, and ALL
      poses of 4 complete (z, beta2) slabs of the 4.48 M-pose list (every cell holds ~276 rows: the reference's FFT
      branch).
cfg4: the full molecules (20 000 + 5 000 atoms), L = 30, Q = 100, all poses of 4 complete (z, beta1, beta2) cells of the
      8.96 M-pose list.
Both files carry the reference's coefficient tables (so the GPU test feeds the scoring path the very numbers the
reference saw), its scores, and the scores of the SENSITIVITY build (same sources, FMA contraction on: oracle/Makefile)
— the reference's own answer to a 1e-16 perturbation of its arithmetic, the noise floor parity is read against.
"""
import os
import sys
import time
from multiprocessing import Pool

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, REPO)
import refso  # noqa: E402
import bench  # noqa: E402
from golden import proto_cross_terms as proto  # noqa: E402


def native_cross(idx, coefA, coefB, q, zv, L):
    A = coefA[..., 0] + 1j * coefA[..., 1]
    B = coefB[..., 0] + 1j * coefB[..., 1]
    return proto.cross_terms(idx, A, B, q, zv, L, proto.reference_tables(L, q, zv))


def _score(job):
    idx, small, so = job
    t = time.time()
    r = refso.scores(idx, small["coefA"], small["coefB"], small["a"], small["scal"], small["qvals"], small["zvals"],
                     small["L"], so=so)
    return r, time.time() - t


def main():
    which = sys.argv[1]
    if which == "cfg3":
        bench.WORKLOAD = "cfg3_3k+1.5k_L15_Q50_70kx64z"
    else:
        bench.WORKLOAD = "cfg4_20k+5k_L30_Q100_70kx128z"
    t = time.time()
    w = bench.build_inputs(refso.expand, refso.opt_params, native_cross, 0, None, None)
    print("inputs: %.1fs, %d poses" % (time.time() - t, len(w["index"])))
    L = w["L"]
    nb, N = L + 1, 2 * L + 1
    if which == "cfg3":
        rows = bench.sample_slabs(w["index"], L, 4, seed=11)
        zsel = None
    else:
        cell = w["index"].astype(np.int64) // N ** 3
        u, cnt = np.unique(cell, return_counts=True)
        rng = np.random.default_rng(11)
        ok = u[(cnt >= 60) & (cnt <= 110)]
        pick = rng.choice(ok, size=4, replace=False)
        rows = [np.flatnonzero(cell == c) for c in pick]
        zsel = None
    # the z table is trimmed to the z steps in use and the z digit renumbered, so the 32-bit entry of the reference
    # serves L = 30 too (SURVEY header note 6); zvals keep their values
    allrows = np.concatenate(rows)
    idx64 = w["index"][allrows].astype(np.int64)
    per_z = nb * nb * N ** 3
    zdig = idx64 // per_z
    zu = np.unique(zdig)
    remap = np.searchsorted(zu, zdig)
    idx_small = (remap * per_z + idx64 % per_z)
    assert idx_small.max() < 2 ** 31
    idx_small = idx_small.astype(np.int32)
    zvals = w["zvals"][zu]
    small = dict(coefA=w["coefA"], coefB=w["coefB"], a=w["a"], scal=w["scal"], qvals=w["qvals"], zvals=zvals, L=L)
    # one process per (slab or cell) and build
    bounds = np.cumsum([0] + [len(r) for r in rows])
    jobs = []
    for so in (None, refso.SENS_SO):
        for i in range(len(rows)):
            jobs.append((idx_small[bounds[i]:bounds[i + 1]], small, so))
    with Pool(min(8, len(jobs))) as pool:
        res = pool.map(_score, jobs)
    k = len(rows)
    ref = [np.concatenate([res[i][0][j] for i in range(k)]) for j in range(3)]
    sens = [np.concatenate([res[k + i][0][j] for i in range(k)]) for j in range(3)]
    print("reference: %s s per part" % ["%.0f" % r[1] for r in res])
    d = [np.abs(sens[0] / ref[0] - 1).max(), np.abs(sens[1] / ref[1] - 1).max(), np.abs(sens[2] - ref[2]).max()]
    print("rows %d; reference vs its FMA build: dchi %.2e dc1 %.2e dc2(abs) %.2e" % (len(idx_small), *d))
    out = dict(L=L, qvals=w["qvals"], zvals=zvals, index=idx_small, a=w["a"], scal=w["scal"],
               coefA=w["coefA"], coefB=w["coefB"], scores=ref[0], c1=ref[1], c2=ref[2],
               sens_scores=sens[0], sens_c1=sens[1], sens_c2=sens[2],
               full_index=w["index"][allrows], full_z=zu, workload=bench.WORKLOAD)
    name = "golden_cfg3_slabs.npz" if which == "cfg3" else "golden_cfg4_cells.npz"
    np.savez_compressed(os.path.join(HERE, name), **out)
    print("wrote", name, os.path.getsize(os.path.join(HERE, name)))


if __name__ == "__main__":
    main()
