"""Regenerates the committed fixtures from the reference's own test data and the reference's own code.

Run in the build container only (needs /root/reference and oracle/_ref/libsxsref.so):

    python tests/golden/make_golden.py [--full]

Outputs (tests/golden/):
  golden_4g9s.npz   inputs of the reference's four Check tests (tests/saxs_test.c) as arrays — atoms of
                    4g9s_r_native / 4g9s_l_moved / 4g9s_native_dimer after prm assignment and centring,
                    SASA fractions, the q grid, ref_saxs, the snapped grid indices of the Euler rows used —
                    plus the reference's goldens (ref_spf, ref_profile, ref_fitted_profile, ref_chi) and the
                    full-precision outputs of the compiled reference on the same inputs.
  golden_real70k.npz (--full) all 70 000 rows of tests/data/euler_coords.000.00 scored by the reference.
The two prm files next to this script are verbatim parameter data from /root/reference/prms.
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import refso  # noqa: E402

D = "/root/reference/tests/data/"
PRM = D + "atoms.0.0.6.prm.ms.3cap+0.5ace.Hr0rec"
MAP = D + "pdb_formfactor_mapping_clean.prm"
L = 15
M_PI = 3.14159265358  # src/define.h:13-15


def c_round(x):
    return np.sign(x) * np.floor(np.abs(x) + 0.5)


def snap(rows, zvals):
    """tools/correlate.c:214-251 / tests/saxs_test.c:283-317"""
    nb, N = L + 1, 2 * L + 1
    b_step = M_PI / L
    a_step = 2.0 * M_PI / N
    idx, ft, order = [], [], []
    for ln, r in enumerate(rows):
        z = r[1]
        for j, zv in enumerate(zvals):
            if zv > z - 0.001 and zv < z + 0.001:
                a2 = 2 * M_PI - r[4]
                g2 = 2 * M_PI - r[6]
                t = j * nb
                t = (t + int(c_round(r[2] / b_step))) * nb
                t = (t + int(c_round(r[5] / b_step))) * N
                t = (t + int(c_round(a2 / a_step))) * N
                t = (t + int(c_round(r[3] / a_step))) * N
                t = t + int(c_round(g2 / a_step))
                idx.append(t)
                ft.append(int(r[0]))
                order.append(ln)
    return np.array(idx, dtype=np.int32), np.array(ft, dtype=np.int32), np.array(order, dtype=np.int32)


def names(lst):
    return np.array(lst, dtype="S8")


def main():
    q = refso.mkarray(0.0, 0.5, 50)
    out = dict(qvals=q, L=L)
    mols = {}
    for key, fn, centre in (("rec", "4g9s_r_native.pdb", 1), ("lig", "4g9s_l_moved.pdb", 2),
                            ("dimer", "4g9s_native_dimer.pdb", 1)):
        m = refso.load_pdb(D + fn, PRM, centre)
        coef, rm, sa = refso.expand(MAP, m["xyz"], m["res"], m["atm"], m["radius"], q, L, water_mode=2)
        mols[key] = (m, coef, rm, sa)
        out.update({key + "_xyz": m["xyz"], key + "_radius": m["radius"], key + "_res": names(m["res"]),
                    key + "_atm": names(m["atm"]), key + "_sa": sa, key + "_coef": coef, key + "_rm": rm,
                    key + "_shift": m["shift"]})
        ff, bad = refso.form_factors(MAP, m["res"], m["atm"])
        assert bad == 0
        out[key + "_ff"] = ff
    # goldens of the reference's tests
    lines = open(D + "ref_spf").read().split("\n")
    out["ref_spf_header"] = np.array([float(x) for x in lines[0].split()])
    out["ref_spf"] = np.array([[float(x) for x in l.split()] for l in lines[1:] if l.strip()]).reshape(50, 256, 6)
    out["ref_profile"] = np.loadtxt(D + "ref_profile")
    out["ref_fitted_profile"] = np.loadtxt(D + "ref_fitted_profile")
    out["ref_chi"] = np.loadtxt(D + "ref_chi")
    # reference outputs at full precision
    rec, A, rmA, _ = mols["rec"]
    lig, B, rmB, _ = mols["lig"]
    pin, perr = refso.profile_from_spf(A, L, rmA, q, 1.0, 1.0)
    out["profile_in"] = pin
    eq, ei, ee = refso.profile_read(D + "ref_saxs")
    out.update(exp_q=eq, exp_in=ei, exp_err=ee)
    nA, nB = len(rec["res"]), len(lig["res"])
    mean_r = (rmA * nA + rmB * nB) / (nA + nB)
    a, scal = refso.opt_params(eq, ei, ee, q, mean_r)
    out.update(a=a, scal=scal)
    dim, Dm, rmD, _ = mols["dimer"]
    a_d, scal_d = refso.opt_params(eq, ei, ee, q, rmD)
    fi, fe, fo = refso.fitted_profile(Dm, L, a_d, scal_d, q)
    out.update(a_dimer=a_d, scal_dimer=scal_d, fitted_in=fi, fitted_out3=fo)

    rows = np.loadtxt(D + "euler_coords.000.00")
    # (1) the reference's own test: z = 40 only (tests/saxs_test.c:187-189)
    idx, ft, _ = snap(rows, [40.0])
    t = time.time()
    s, c1, c2 = refso.scores(idx, A, B, a, scal, q, [40.0], L)
    print("z=40: %d rows, %.1fs" % (len(idx), time.time() - t))
    out.update(z40_index=idx, z40_ft=ft, z40_scores=s, z40_c1=c1, z40_c2=c2)
    # (2) six z steps incl. cells that take the reference's FFT branch (>= 30 rows)
    zv = [20.0, 21.0, 22.0, 40.0, 41.0, 42.0]
    idx, ft, order = snap(rows, zv)
    t = time.time()
    s, c1, c2 = refso.scores(idx, A, B, a, scal, q, zv, L)
    print("6 z: %d rows, %.1fs" % (len(idx), time.time() - t))
    out.update(z6_zvals=np.array(zv), z6_index=idx, z6_ft=ft, z6_order=order, z6_scores=s, z6_c1=c1, z6_c2=c2)
    np.savez_compressed(os.path.join(HERE, "golden_4g9s.npz"), **out)

    if "--full" in sys.argv:
        from multiprocessing import Pool
        zall = [float(z) for z in range(1, 81)]  # tools/correlate.c:39-41
        idx, ft, order = snap(rows, zall)
        nb, N = L + 1, 2 * L + 1
        zdig = idx // (nb * nb * N ** 3)
        jobs = [(z, idx[zdig == z], A, B, a, scal, q, zall) for z in np.unique(zdig)]
        t = time.time()
        with Pool(8) as pool:
            res = pool.map(_one_z, jobs)
        s = np.zeros(len(idx)); c1 = np.zeros(len(idx)); c2 = np.zeros(len(idx))
        for (z, r) in res:
            m = zdig == z
            s[m], c1[m], c2[m] = r
        print("full list: %d rows, %.1fs" % (len(idx), time.time() - t))
        np.savez_compressed(os.path.join(HERE, "golden_real70k.npz"), index=idx, ft=ft, order=order,
                            zvals=np.array(zall), scores=s, c1=c1, c2=c2)


def _one_z(job):
    z, idx, A, B, a, scal, q, zall = job
    return z, refso.scores(idx, A, B, a, scal, q, zall, L)


if __name__ == "__main__":
    main()
