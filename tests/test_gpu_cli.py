"""CLI contract: our `correlate` / `single_saxs` binaries against the reference's own tools (compiled unmodified
into oracle/_ref) on the same files — argument list, Euler side file, output rows and their order."""
import os
import subprocess

import numpy as np
import pytest

import refso

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
GOLD = os.path.join(HERE, "golden")
MAP = os.path.join(GOLD, "pdb_formfactor_mapping_clean.prm")
PRM = os.path.join(GOLD, "atoms.prm")
BIN = os.path.join(REPO, "libfmftsaxs_b200", "bin")
REF = os.path.join(REPO, "oracle", "_ref")

pytestmark = pytest.mark.gpu


def write_pdb(path, res, atm, xyz):
    with open(path, "w") as f:
        f.write("REMARK fixture\n")
        for i, (r, a, x) in enumerate(zip(res, atm, xyz)):
            f.write("ATOM  %5d %-4s %-4s %4d    %8.3f%8.3f%8.3f  1.00  0.00\n" % (i + 1, a.decode(), r.decode().strip(), i // 10 + 1, x[0], x[1], x[2]))
        f.write("END\n")


@pytest.fixture(scope="module")
def files(tmp_path_factory):
    G = np.load(os.path.join(GOLD, "golden_4g9s.npz"))
    d = tmp_path_factory.mktemp("cli")
    write_pdb(d / "rec.pdb", G["rec_res"], G["rec_atm"], G["rec_xyz"] - G["rec_shift"])
    write_pdb(d / "lig.pdb", G["lig_res"], G["lig_atm"], G["lig_xyz"] - G["lig_shift"])
    with open(d / "exp.dat", "w") as f:
        for q, i, e in zip(G["exp_q"], G["exp_in"], G["exp_err"]):
            f.write(" %.6e  %.6e  %.6e\n" % (q, i, e))
    rng = np.random.default_rng(21)
    nrot, nrow = 40, 48
    with open(d / "rot.prm", "w") as f:
        for k in range(nrot):
            qn = rng.normal(size=4)
            qn /= np.linalg.norm(qn)
            w, x, y, z = qn
            R = [1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w), 2 * (x * y + z * w),
                 1 - 2 * (x * x + z * z), 2 * (y * z - x * w), 2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]
            f.write("%d " % k + " ".join("%.9f" % v for v in R) + "\n")
    # ligand centre relative to receptor centre in the input frames = -(com - coe) with the stored negated shifts
    ref_lig = -(G["lig_shift"] - G["rec_shift"])
    with open(d / "ft.000", "w") as f:
        for k in range(nrow):
            u = rng.normal(size=3)
            u /= np.linalg.norm(u)
            dist = rng.choice([38.0, 39.0, 95.0]) + rng.uniform(-0.3, 0.3)   # 95 A is off the 1..80 table: row dropped
            t = dist * u - ref_lig
            f.write("%d %.3f %.3f %.3f 0.0 0.0 0.0 0.0 0.0 0.0\n" % (rng.integers(0, nrot), t[0], t[1], t[2]))
    return d


@pytest.mark.skipif(not os.path.exists(os.path.join(REF, "ref_correlate")), reason="compiled reference tools not present")
def test_correlate_cli_matches_reference_tool(files):
    d = files
    common = [MAP, PRM, str(d / "ft.000"), str(d / "rot.prm"), str(d / "rec.pdb"), str(d / "lig.pdb"), str(d / "exp.dat"), "15"]
    r1 = subprocess.run([os.path.join(BIN, "correlate")] + common + [str(d / "eul_ours"), str(d / "out_ours")],
                        stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert r1.returncode == 0, r1.stdout[-2000:]
    assert "Time passed:" in r1.stdout and "Correlation finished" in r1.stdout
    r2 = subprocess.run([os.path.join(REF, "ref_correlate")] + common + [str(d / "eul_ref"), str(d / "out_ref")],
                        stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    assert r2.returncode == 0, r2.stdout[-2000:]
    assert open(d / "eul_ours").read() == open(d / "eul_ref").read()
    ours = [l.split("\t") for l in open(d / "out_ours").read().splitlines()]
    ref = [l.split("\t") for l in open(d / "out_ref").read().splitlines()]
    assert len(ours) == len(ref) and 0 < len(ours) < 48          # the 95 A rows are dropped by both
    for a, b in zip(ours, ref):
        assert a[0].strip() == b[0].strip() and a[1] == b[1]       # serial and ft id, same order
        for x, y in zip(a[2:], b[2:]):
            assert abs(float(x) - float(y)) <= 1.001e-3            # "%.3lf" columns


@pytest.mark.skipif(not os.path.exists(os.path.join(REF, "ref_single_saxs")), reason="compiled reference tools not present")
def test_single_saxs_cli_matches_reference_tool(files):
    d = files
    for L in ("10", "15", "20"):                                   # examples/run_single_saxs.sh sweep of BASELINE config 2
        common = [MAP, PRM, str(d / "rec.pdb"), str(d / "lig.pdb"), "1.0", "0.0", L]
        subprocess.run([os.path.join(BIN, "single_saxs")] + common + [str(d / "p_ours")], check=True, stdout=subprocess.DEVNULL)
        subprocess.run([os.path.join(REF, "ref_single_saxs")] + common + [str(d / "p_ref")], check=True, stdout=subprocess.DEVNULL)
        a, b = np.loadtxt(d / "p_ours"), np.loadtxt(d / "p_ref")
        assert a.shape == b.shape == (50, 3)
        assert np.max(np.abs(a[:, 1] / b[:, 1] - 1)) < 1e-6


@pytest.mark.skipif(not os.path.exists(os.path.join(REF, "ref_score_ft_naive")), reason="compiled reference tools not present")
def test_score_ft_naive_cli_matches_reference_tool_and_the_fft_path(files):
    """SURVEY 8f-3, tools/score_ft_naive.c: the real-space validator (pose built atom by atom, K1 + self terms + K4 per
    pose) gives the reference tool's rows, and agrees with `correlate` as far as the truncation at L = 15 allows"""
    d = files
    rows = open(d / "ft.000").read().splitlines()[:10]
    with open(d / "ft.naive", "w") as f:
        f.write("\n".join(rows) + "\n")
    common = [MAP, PRM, str(d / "ft.naive"), str(d / "rot.prm"), str(d / "rec.pdb"), str(d / "lig.pdb"), str(d / "exp.dat"), "15"]
    for exe, out in ((os.path.join(BIN, "score_ft_naive"), "naive_ours"), (os.path.join(REF, "ref_score_ft_naive"), "naive_ref")):
        wd = d / ("wd_" + out)
        os.makedirs(wd, exist_ok=True)
        r = subprocess.run([exe] + common + [str(d / out)], cwd=wd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                           text=True, timeout=900)
        assert r.returncode == 0, r.stdout[-2000:]
    assert open(d / "wd_naive_ours" / "euler_list").read() == open(d / "wd_naive_ref" / "euler_list").read()
    ours, ref = np.loadtxt(d / "naive_ours"), np.loadtxt(d / "naive_ref")
    assert ours.shape == ref.shape == (10, 4)
    assert np.array_equal(ours[:, 0], ref[:, 0])
    assert np.max(np.abs(ours[:, 1:] - ref[:, 1:])) <= 1.001e-3      # "%.3f" columns
    # against the FFT path on the rows that lie on its z table
    subprocess.run([os.path.join(BIN, "correlate")] + common + [str(d / "eul_n"), str(d / "out_n")], check=True,
                   stdout=subprocess.DEVNULL, timeout=600)
    fft = {int(l.split("\t")[0]): [float(x) for x in l.split("\t")[2:]] for l in open(d / "out_n").read().splitlines()}
    assert len(fft) >= 3
    dev = np.array([[ours[i, 1] / fft[i][0] - 1, ours[i, 2] - fft[i][1], ours[i, 3] - fft[i][2]] for i in fft])
    print("naive vs FFT path on %d rows: max rel dchi %.3g, max dc1 %.3g, max dc2 %.3g" % (len(fft), np.abs(dev[:, 0]).max(),
          np.abs(dev[:, 1]).max(), np.abs(dev[:, 2]).max()))
    # different hydration weights (joined vs separate molecules) and different truncations: same physics, not same digits
    assert np.abs(dev[:, 0]).max() < 0.25


@pytest.mark.skipif(not os.path.exists(os.path.join(REF, "dropin_correlate")), reason="drop-in tools not built (need /root/reference at build time)")
def test_reference_tool_sources_run_on_the_product_library(files):
    """Drop-in proof: the reference's OWN tools/correlate.c and tools/single_saxs.c, compiled unmodified against include/
    and linked to libfmftsaxs.so (oracle/Makefile `dropin`), give the rows of the all-reference build"""
    d = files
    common = [MAP, PRM, str(d / "ft.000"), str(d / "rot.prm"), str(d / "rec.pdb"), str(d / "lig.pdb"), str(d / "exp.dat"), "15"]
    r = subprocess.run([os.path.join(REF, "dropin_correlate")] + common + [str(d / "eul_drop"), str(d / "out_drop")],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:]
    if not os.path.exists(d / "out_ref"):
        subprocess.run([os.path.join(REF, "ref_correlate")] + common + [str(d / "eul_ref"), str(d / "out_ref")], check=True,
                       stdout=subprocess.DEVNULL, timeout=900)
    assert open(d / "eul_drop").read() == open(d / "eul_ref").read()
    drop = [l.split("\t") for l in open(d / "out_drop").read().splitlines()]
    ref = [l.split("\t") for l in open(d / "out_ref").read().splitlines()]
    assert len(drop) == len(ref) > 0
    for a, b in zip(drop, ref):
        assert a[0].strip() == b[0].strip() and a[1] == b[1]
        for x, y in zip(a[2:], b[2:]):
            assert abs(float(x) - float(y)) <= 1.001e-3
    s_common = [MAP, PRM, str(d / "rec.pdb"), str(d / "lig.pdb"), "1.0", "0.0", "15"]
    subprocess.run([os.path.join(REF, "dropin_single_saxs")] + s_common + [str(d / "p_drop")], check=True, stdout=subprocess.DEVNULL)
    subprocess.run([os.path.join(REF, "ref_single_saxs")] + s_common + [str(d / "p_ref2")], check=True, stdout=subprocess.DEVNULL)
    a, b = np.loadtxt(d / "p_drop"), np.loadtxt(d / "p_ref2")
    assert a.shape == b.shape == (50, 3) and np.max(np.abs(a[:, 1] / b[:, 1] - 1)) < 1e-6


@pytest.mark.skipif(not os.path.exists(os.path.join(REF, "ref_correlate")), reason="compiled reference tools not present")
def test_correlate_cli_keeps_64_bit_indices_where_int_overflows(files):
    """L = 20 with the 80-step z table needs more than 31 bits for the flat index; the tool then keeps 64-bit indices
    (the reference's `int` packing is still right for z < 70, which is where these rows are): same rows as the reference"""
    d = files
    rows = open(d / "ft.000").read().splitlines()[:16]
    with open(d / "ft.l20", "w") as f:
        f.write("\n".join(rows) + "\n")
    common = [MAP, PRM, str(d / "ft.l20"), str(d / "rot.prm"), str(d / "rec.pdb"), str(d / "lig.pdb"), str(d / "exp.dat"), "20"]
    subprocess.run([os.path.join(BIN, "correlate")] + common + [str(d / "eul20_ours"), str(d / "out20_ours")], check=True,
                   stdout=subprocess.DEVNULL, timeout=600)
    subprocess.run([os.path.join(REF, "ref_correlate")] + common + [str(d / "eul20_ref"), str(d / "out20_ref")], check=True,
                   stdout=subprocess.DEVNULL, timeout=1500)
    assert open(d / "eul20_ours").read() == open(d / "eul20_ref").read()
    ours = [l.split("\t") for l in open(d / "out20_ours").read().splitlines()]
    ref = [l.split("\t") for l in open(d / "out20_ref").read().splitlines()]
    assert len(ours) == len(ref) and len(ours) >= 5
    for a, b in zip(ours, ref):
        assert a[0].strip() == b[0].strip() and a[1] == b[1]
        for x, y in zip(a[2:], b[2:]):
            assert abs(float(x) - float(y)) <= 1.001e-3


@pytest.mark.skipif(not os.path.exists(os.path.join(REF, "ref_correlate")), reason="compiled reference tools not present")
def test_config1_run_correlate_on_dimer2(tmp_path):
    """BASELINE config 1, examples/run_correlate.sh:40-48 on dimer2 (1A2K): single_saxs -> reference profile, the three
    ft files concatenated, all 210 000 rows through `correlate`; the sub-list {z = 33, u_z > 0.85} row by row against the
    reference tool, and the same poses inside the full run carry the same numbers (scripts/config1.py)"""
    import sys
    sys.path.insert(0, os.path.join(REPO, "scripts"))
    import config1
    res = config1.main(str(tmp_path / "c1"))
    print(res)
    assert res["rows_scored"] == 210000 and res["subset_rows"] > 300
    assert res["euler_file_identical"] and res["subset_max_abs_diff_printed"] <= 1.001e-3
