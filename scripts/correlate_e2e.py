#!/usr/bin/env python
"""End-to-end wall time of the `correlate` tool (file parsing + expansion + Euler text round trip + scoring +
output included) on the 4G9S test molecules with a synthetic ft file of N rows over 70 000 rotations.

    python scripts/correlate_e2e.py [N=2000000]
"""
import os
import subprocess
import sys
import time

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "tests"))
GOLD = os.path.join(REPO, "tests", "golden")


def main():
    n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 2000000
    from test_gpu_cli import write_pdb
    G = np.load(os.path.join(GOLD, "golden_4g9s.npz"))
    d = os.path.join(REPO, "gpurun_out", "e2e")
    os.makedirs(d, exist_ok=True)
    write_pdb(os.path.join(d, "rec.pdb"), G["rec_res"], G["rec_atm"], G["rec_xyz"] - G["rec_shift"])
    write_pdb(os.path.join(d, "lig.pdb"), G["lig_res"], G["lig_atm"], G["lig_xyz"] - G["lig_shift"])
    with open(os.path.join(d, "exp.dat"), "w") as f:
        for q, i, e in zip(G["exp_q"], G["exp_in"], G["exp_err"]):
            f.write(" %.6e  %.6e  %.6e\n" % (q, i, e))
    rng = np.random.default_rng(5)
    nrot = 70000
    qn = rng.normal(size=(nrot, 4))
    qn /= np.linalg.norm(qn, axis=1, keepdims=True)
    w, x, y, z = qn.T
    R = np.stack([1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w), 2 * (x * y + z * w),
                  1 - 2 * (x * x + z * z), 2 * (y * z - x * w), 2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)], 1)
    np.savetxt(os.path.join(d, "rot.prm"), np.column_stack([np.arange(nrot), R]), fmt="%d" + " %.9f" * 9)
    ref_lig = -(G["lig_shift"] - G["rec_shift"])
    u = rng.normal(size=(n, 3))
    u /= np.linalg.norm(u, axis=1, keepdims=True)
    dist = rng.integers(20, 46, n) + rng.uniform(-0.3, 0.3, n)
    t = dist[:, None] * u - ref_lig
    t0 = time.time()
    np.savetxt(os.path.join(d, "ft.000"), np.column_stack([rng.integers(0, nrot, n), t, np.zeros((n, 6))]),
               fmt="%d %.3f %.3f %.3f" + " %.1f" * 6)
    print("ft file: %d rows, %.1f MB, written in %.1f s" % (n, os.path.getsize(os.path.join(d, "ft.000")) / 1e6, time.time() - t0), flush=True)
    cmd = [os.path.join(REPO, "libfmftsaxs_b200", "bin", "correlate"), os.path.join(GOLD, "pdb_formfactor_mapping_clean.prm"),
           os.path.join(GOLD, "atoms.prm"), os.path.join(d, "ft.000"), os.path.join(d, "rot.prm"), os.path.join(d, "rec.pdb"),
           os.path.join(d, "lig.pdb"), os.path.join(d, "exp.dat"), "15", os.path.join(d, "euler.out"), os.path.join(d, "scores.out")]
    for rep in range(2):
        t0 = time.time()
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=dict(os.environ, SXS_TIMING="1"))
        dt = time.time() - t0
        assert r.returncode == 0, r.stdout[-2000:]
        rows = sum(1 for _ in open(os.path.join(d, "scores.out")))
        cpu = [l for l in r.stdout.splitlines() if l.startswith("Time passed")]
        print("run %d: correlate wall %.2f s, %d scored rows, %.0f rows/s end to end; %s (CPU clock of the tool)" %
              (rep, dt, rows, rows / dt, cpu[-1] if cpu else ""), flush=True)
        print("\n".join(l for l in r.stdout.splitlines() if l.startswith("[phase]")), flush=True)
    for f in ("ft.000", "euler.out", "scores.out"):
        os.remove(os.path.join(d, f))


if __name__ == "__main__":
    main()
