#!/usr/bin/env python
"""BASELINE config 1 end to end: examples/run_correlate.sh on dimer2 (1A2K) — single_saxs writes the reference profile
(c1 = c2 = 1.0, L = 15), the three ft files are concatenated, `correlate` scores all 210 000 rows — with the seeded
stand-ins of the missing ft / rotation files (libfmftsaxs_b200/workload.py:make_config1).

    python scripts/config1.py [workdir]        prints one JSON line (wall times, rows/s, parity on the sampled subset)

Parity: the reference tool (oracle/_ref/ref_correlate, the reference's unmodified sources) on all 210 000 rows costs
~30 min on 16 cores; it is run on the sub-list {z = 33, u_z > 0.85} (~600 rows, ~50 cells) and compared row by row
(serial, ft id, three printed decimals); our tool's rows for the same poses inside the full run are then required to
carry the same numbers.
"""
import json
import os
import subprocess
import sys
import time

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))
GOLD = os.path.join(REPO, "tests", "golden")
MAP = os.path.join(GOLD, "pdb_formfactor_mapping_clean.prm")
PRM = os.path.join(GOLD, "atoms.prm")
BIN = os.path.join(REPO, "libfmftsaxs_b200", "bin")
REF = os.path.join(REPO, "oracle", "_ref")


def run(cmd, **kw):
    t = time.time()
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, **kw)
    if r.returncode != 0:
        raise RuntimeError("%s failed:\n%s" % (cmd[0], r.stdout[-2000:]))
    return time.time() - t, r.stdout


def prepare(d, nrot=70000):
    from libfmftsaxs_b200 import workload as wl
    from test_gpu_cli import write_pdb
    G = np.load(os.path.join(GOLD, "golden_dimer2.npz"))
    os.makedirs(d, exist_ok=True)
    write_pdb(os.path.join(d, "r_u_nmin.pdb"), G["rec_res"], G["rec_atm"], G["rec_xyz"])
    write_pdb(os.path.join(d, "l_u_nmin.pdb"), G["lig_res"], G["lig_atm"], G["lig_xyz"])
    rot, files = wl.make_config1(G["rec_xyz"], G["lig_xyz"], nrot=nrot)
    wl.write_rm_file(os.path.join(d, "rot70k.prm"), rot)
    for k, f in enumerate(files):
        wl.write_ft_file(os.path.join(d, "ft.%03d.00" % k), f["rot"], f["t"])
    with open(os.path.join(d, "ft_combo"), "w") as out:              # `cat ft1 ft2 ft3 > ft_combo`
        for k in range(len(files)):
            out.write(open(os.path.join(d, "ft.%03d.00" % k)).read())
    # the sub-list the CPU reference can afford
    sel = [np.flatnonzero((f["z"] == 33) & (f["u"][:, 2] > 0.85)) for f in files]
    with open(os.path.join(d, "ft_subset"), "w") as out:
        for k, s in enumerate(sel):
            lines = open(os.path.join(d, "ft.%03d.00" % k)).read().splitlines()
            for i in s:
                out.write(lines[i] + "\n")
    combo_rows = np.concatenate([s + k * nrot for k, s in enumerate(sel)])   # their serial numbers in the full run
    return files, combo_rows


def main(d=None, with_reference=True, nrot=70000):
    d = d or os.path.join(REPO, "gpurun_out", "config1")
    t0 = time.time()
    files, combo_rows = prepare(d, nrot)
    t_prep = time.time() - t0
    rec, lig = os.path.join(d, "r_u_nmin.pdb"), os.path.join(d, "l_u_nmin.pdb")
    exp = os.path.join(d, "ref_saxs_profile")
    t_single, _ = run([os.path.join(BIN, "single_saxs"), MAP, PRM, rec, lig, "1.0", "1.0", "15", exp])
    common = [MAP, PRM, None, os.path.join(d, "rot70k.prm"), rec, lig, exp, "15"]
    full = list(common); full[2] = os.path.join(d, "ft_combo")
    run([os.path.join(BIN, "correlate")] + full + [os.path.join(d, "euler_list"), os.path.join(d, "chi_scores")])   # warm-up (page cache, CUDA context)
    t_corr, out = run([os.path.join(BIN, "correlate")] + full + [os.path.join(d, "euler_list"), os.path.join(d, "chi_scores")])
    rows = open(os.path.join(d, "chi_scores")).read().splitlines()
    n_rows = len(rows)
    res = {"config": "BASELINE config 1: examples/run_correlate.sh, dimer2 (1A2K), %d ft rows in %d files, L = 15" % (3 * nrot, 3),
           "rows_scored": n_rows, "single_saxs_s": t_single, "correlate_wall_s": t_corr,
           "rows_per_s_end_to_end": n_rows / t_corr, "prepare_inputs_s": t_prep,
           "tool_clock": [l for l in out.splitlines() if l.startswith("Time passed")][-1:]}
    sub = list(common); sub[2] = os.path.join(d, "ft_subset")
    run([os.path.join(BIN, "correlate")] + sub + [os.path.join(d, "euler_sub_ours"), os.path.join(d, "chi_sub_ours")])
    ours_sub = [l.split("\t") for l in open(os.path.join(d, "chi_sub_ours")).read().splitlines()]
    # the same poses inside the full run: same three printed numbers
    by_serial = {int(l.split("\t")[0]): l.split("\t")[2:] for l in rows}
    assert len(ours_sub) == len(combo_rows)
    for a, serial in zip(ours_sub, combo_rows):
        assert by_serial[int(serial)] == a[2:], (serial, by_serial[int(serial)], a)
    res["subset_rows"] = len(ours_sub)
    if with_reference and os.path.exists(os.path.join(REF, "ref_correlate")):
        t_ref, _ = run([os.path.join(REF, "ref_correlate")] + sub + [os.path.join(d, "euler_sub_ref"), os.path.join(d, "chi_sub_ref")])
        assert open(os.path.join(d, "euler_sub_ours")).read() == open(os.path.join(d, "euler_sub_ref")).read()
        ref_sub = [l.split("\t") for l in open(os.path.join(d, "chi_sub_ref")).read().splitlines()]
        assert len(ref_sub) == len(ours_sub)
        worst = 0.0
        for a, b in zip(ours_sub, ref_sub):
            assert a[0].strip() == b[0].strip() and a[1] == b[1]
            worst = max(worst, max(abs(float(x) - float(y)) for x, y in zip(a[2:], b[2:])))
        assert worst <= 1.001e-3
        res.update(reference_subset_s=t_ref, reference_rows_per_s_one_core=len(ref_sub) / t_ref,
                   subset_max_abs_diff_printed=worst, euler_file_identical=True)
    for f in ("ft_combo", "euler_list"):
        os.remove(os.path.join(d, f))
    return res


if __name__ == "__main__":
    print(json.dumps(main(sys.argv[1] if len(sys.argv) > 1 else None)))
