"""Seeded synthetic workloads of the sizes BASELINE.json names, numpy only.

A workload is everything `correlate` would have in memory right before sxs_compute_saxs_scores:
two atom groups (names, coordinates, radii, SASA fractions), the q grid, the z table, the flat grid
indices of the pose list, and an "experimental" curve.  Both bench arms (GPU product, CPU reference)
consume the same arrays; neither implementation is used to make them.
"""
import os

import numpy as np

M_PI = 3.14159265358  # src/define.h:13-15 — the grid steps use the truncated constant

PKG = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(os.path.dirname(PKG), "tests", "golden")
MAP_PATH = os.path.join(GOLD, "pdb_formfactor_mapping_clean.prm")
PRM_PATH = os.path.join(GOLD, "atoms.prm")

CONFIGS = {
    # name: (rec atoms, lig atoms, L, qnum, rotations, z table)
    "cfg3_3k+1.5k_L15_Q50_70kx64z": dict(n_rec=3000, n_lig=1500, L=15, qnum=50, nrot=70000,
                                         zvals=np.arange(17.0, 81.0, 1.0)),
    "cfg4_20k+5k_L30_Q100_70kx128z": dict(n_rec=20000, n_lig=5000, L=30, qnum=100, nrot=70000,
                                          zvals=np.arange(16.5, 80.01, 0.5)),
}


def _atom_table():
    """(res, atom, radius) rows present in both parameter files with a type the reference can parse"""
    known = {"H", "HE", "C", "N", "O", "NE", "SOD+", "MG2+", "P", "S", "K", "CAL2+", "FE2+", "ZN2+", "SE", "AU",
             "CH", "CH2", "CH3", "NH", "NH2", "NH3", "OH", "SH"}
    mapping = {}
    for line in open(MAP_PATH):
        if line.startswith("#") or not line.strip():
            continue
        f = line.split()
        if len(f) >= 3 and f[2] in known:
            mapping[(f[0], f[1])] = f[2]
    rows = []
    seen = set()
    for line in open(PRM_PATH):
        f = line.split()
        if len(f) == 7 and f[0] == "atom":
            key = (f[2], f[3])
            if key in mapping and key not in seen and len(f[2]) <= 4 and len(f[3]) <= 4:
                seen.add(key)
                rows.append((f[2], f[3], float(f[5])))
    return rows


def make_molecule(natoms, seed):
    rng = np.random.default_rng(seed)
    table = _atom_table()
    pick = rng.integers(0, len(table), natoms)
    res = [table[i][0] for i in pick]
    atm = [table[i][1] for i in pick]
    radius = np.array([table[i][2] for i in pick])
    R = 1.35 * natoms ** (1.0 / 3.0)
    xyz = np.zeros((natoms, 3))
    k = 0
    while k < natoms:
        p = rng.uniform(-R, R, (2 * natoms, 3))
        ok = (np.linalg.norm(p, axis=1) <= R) & (np.hypot(p[:, 0], p[:, 1]) > 0.5)
        p = p[ok][: natoms - k]
        xyz[k:k + len(p)] = p
        k += len(p)
    sa = np.where((radius > 0) & (np.linalg.norm(xyz, axis=1) > 0.7 * R), rng.uniform(0, 1, natoms), 0.0)
    return dict(xyz=xyz, res=res, atm=atm, radius=radius, sa=sa)


def _c_round(x):
    return np.sign(x) * np.floor(np.abs(x) + 0.5)


def _random_rotations(n, rng):
    qn = rng.normal(size=(n, 4))
    qn /= np.linalg.norm(qn, axis=1, keepdims=True)
    w, x, y, z = qn.T
    return np.stack([1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w),
                     2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w),
                     2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)], 1).reshape(n, 3, 3)


def make_pose_indices(L, zvals, nrot, seed, dtype=None):
    """Vectorised sxs_ft2euler (src/index.c:38-75) -> 3-decimal text quantisation (src/index.c:114) -> grid
    snapping (tools/correlate.c:219-242): one pose per (rotation, z step), random approach direction."""
    rng = np.random.default_rng(seed)
    nb, N = L + 1, 2 * L + 1
    nz = len(zvals)
    rot = _random_rotations(nrot, rng)
    out = []
    for zi in range(nz):
        u = rng.normal(size=(nrot, 3))
        u /= np.linalg.norm(u, axis=1, keepdims=True)
        bad = np.abs(u[:, 2]) > 0.999
        u[bad] = np.array([0.6, 0.0, 0.8])
        b1 = np.arccos(np.clip(u[:, 2], -1, 1))
        sb1 = np.sin(b1)
        g1 = np.arccos(np.clip(-u[:, 0] / sb1, -1, 1))
        g1 = np.where(u[:, 1] / sb1 < 0, 2 * M_PI - g1, g1)
        # receptor frame rotation R(0, b1, g1) (src/saxs_utils.c:65-79), then R_lig' = R_rec * R
        cb, sb, cg, sg = np.cos(b1), np.sin(b1), np.cos(g1), np.sin(g1)
        rec = np.zeros((nrot, 3, 3))
        rec[:, 0, 0] = cg * cb; rec[:, 1, 0] = sg; rec[:, 2, 0] = -cg * sb
        rec[:, 0, 1] = -sg * cb; rec[:, 1, 1] = cg; rec[:, 2, 1] = sg * sb
        rec[:, 0, 2] = sb; rec[:, 1, 2] = 0.0; rec[:, 2, 2] = cb
        lig = np.einsum("nij,njk->nik", rec, rot)
        b2 = np.arccos(np.clip(lig[:, 2, 2], -1, 1))
        sb2 = np.sin(b2)
        a2 = np.arccos(np.clip(lig[:, 0, 2] / sb2, -1, 1))
        a2 = np.where(lig[:, 1, 2] / sb2 < 0, 2 * M_PI - a2, a2)
        g2 = np.arccos(np.clip(-lig[:, 2, 0] / sb2, -1, 1))
        g2 = np.where(lig[:, 2, 1] / sb2 < 0, 2 * M_PI - g2, g2)
        ang = np.round(np.stack([b1, g1, a2, b2, g2], 1), 3)  # "% .3f"
        b_step, a_step = M_PI / L, 2.0 * M_PI / N
        t = np.full(nrot, zi, dtype=np.int64) * nb
        t = (t + _c_round(ang[:, 0] / b_step).astype(np.int64)) * nb
        t = (t + _c_round(ang[:, 3] / b_step).astype(np.int64)) * N
        t = (t + _c_round((2 * M_PI - ang[:, 2]) / a_step).astype(np.int64)) * N
        t = (t + _c_round(ang[:, 1] / a_step).astype(np.int64)) * N
        t = t + _c_round((2 * M_PI - ang[:, 4]) / a_step).astype(np.int64)
        out.append(t)
    idx = np.concatenate(out)
    if dtype is None:
        dtype = np.int32 if idx.max() < 2 ** 31 else np.int64
    return idx.astype(dtype)


def make_qvals(qnum, qmax=0.5):
    """sxs_mkarray(0, QMAX, qnum): accumulated steps (src/saxs_utils.c:50-63)"""
    q = np.zeros(qnum)
    step = (qmax - 0.0) / (qnum - 1)
    for i in range(1, qnum):
        q[i] = q[i - 1] + step
    return q


def experimental_curve(coefA, coefB, qvals):
    """I(q) of the two molecules far apart at c1 = c2 = 1 (self terms only), 5 % errors"""
    A = coefA[..., 0] + 1j * coefA[..., 1]
    B = coefB[..., 0] + 1j * coefB[..., 1]
    amp = lambda M: M[0] - M[1] + M[2]
    I = (np.abs(amp(A)) ** 2).sum(-1) + (np.abs(amp(B)) ** 2).sum(-1)
    return qvals.copy(), I, 0.05 * I


def mean_radius(rec, lig):
    """tools/correlate.c:125-126"""
    nA, nB = len(rec["radius"]), len(lig["radius"])
    return (rec["radius"].mean() * nA + lig["radius"].mean() * nB) / (nA + nB)


def make(name, seed=20240917, nrot=None, nz=None):
    cfg = dict(CONFIGS[name])
    if nrot is not None:
        cfg["nrot"] = nrot
    zvals = cfg["zvals"] if nz is None else cfg["zvals"][:nz]
    rec = make_molecule(cfg["n_rec"], seed + 1)
    lig = make_molecule(cfg["n_lig"], seed + 2)
    # receptor centred on its centre of extrema, ligand on its centroid (tools/correlate.c:83-105)
    rec["xyz"] -= 0.5 * (rec["xyz"].min(0) + rec["xyz"].max(0))
    lig["xyz"] -= lig["xyz"].mean(0)
    idx = make_pose_indices(cfg["L"], zvals, cfg["nrot"], seed + 3)
    return dict(name=name, L=cfg["L"], qvals=make_qvals(cfg["qnum"]), zvals=np.asarray(zvals, dtype=np.float64),
                rec=rec, lig=lig, index=idx, seed=seed)


# ------------------------------------------------------------------ BASELINE config 1 (examples/run_correlate.sh)

def _splitmix64(seed, n):
    """n 64-bit outputs of the splitmix64 stream started at `seed` (vectorised)"""
    with np.errstate(over="ignore"):
        z = np.uint64(seed) + np.uint64(0x9E3779B97F4A7C15) * np.arange(1, n + 1, dtype=np.uint64)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def _uniform01(seed, n):
    return (_splitmix64(seed, n) >> np.uint64(11)).astype(np.float64) / 9007199254740992.0


def _normals(seed, n):
    u = _uniform01(seed, 2 * n)
    return np.sqrt(-2.0 * np.log(1.0 - u[:n])) * np.cos(2.0 * np.pi * u[n:])


def make_config1(rec_xyz, lig_xyz, nrot=70000, nfiles=3, seed=0x5A170001, z_lo=20, z_hi=45):
    """The inputs examples/run_correlate.sh needs and the reference does not ship (.MISSING_LARGE_BLOBS): a rotation
    set of `nrot` matrices and `nfiles` ft files of `nrot` rows each ("Config 1").

    rotation k   = unit quaternion from a splitmix64 stream (seed 0x5A170001), written as an rm file row
    ft row k     = rotation index k (each rotation once per file, like the real PIPER fixture), translation
                   t = z u - ref_lig with u a seeded unit vector (|u_z| <= 0.999) and integer z ~ U{z_lo..z_hi};
                   ref_lig = ligand centroid - receptor centre of extrema, both in the input frame
    Returns (rot [nrot][9], [ (rot_index, t[3], z, u) per file ]).
    """
    qn = _normals(seed, 4 * nrot).reshape(nrot, 4)
    qn /= np.linalg.norm(qn, axis=1, keepdims=True)
    w, x, y, z = qn.T
    rot = np.stack([1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w),
                    2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w),
                    2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)], 1)
    rec_c = 0.5 * (rec_xyz.min(0) + rec_xyz.max(0))
    ref_lig = lig_xyz.mean(0) - rec_c
    files = []
    for f in range(nfiles):
        s = seed + 7919 * (f + 1)
        u = _normals(s, 3 * nrot).reshape(nrot, 3)
        u /= np.linalg.norm(u, axis=1, keepdims=True)
        bad = np.abs(u[:, 2]) > 0.999
        u[bad] = np.array([0.6, 0.0, 0.8])
        zz = z_lo + (_splitmix64(s + 1, nrot) % np.uint64(z_hi - z_lo + 1)).astype(np.int64)
        t = zz[:, None] * u - ref_lig
        files.append(dict(rot=np.arange(nrot), t=t, z=zz, u=u))
    return rot, files


def write_rm_file(path, rot):
    """one row-major 3x3 per line with a leading index (the layout mol_matrix3_list_from_file takes)"""
    np.savetxt(path, np.column_stack([np.arange(len(rot)), rot]), fmt="%d" + " %.9f" * 9)


def write_ft_file(path, rows_rot, rows_t):
    """ft rows: rotation index, translation, six ignored columns (src/index.c:87-89)"""
    np.savetxt(path, np.column_stack([rows_rot, rows_t, np.zeros((len(rows_rot), 6))]), fmt="%d %.3f %.3f %.3f" + " %.1f" * 6)
