"""Aggregates an `ncu --page source --csv --print-source sass` export of k_fit by device function, using the symbol
table of the cubin inside libfmftsaxs.so (the de-inlined optimiser functions are local symbols of the kernel).

    ncu -i REPORT.ncu-rep --page source --csv --print-source sass > sass.csv
    python scripts/ncu_by_function.py sass.csv [path/to/libfmftsaxs.so]
"""
import csv
import os
import re
import subprocess
import sys
import tempfile


def main():
    sass = sys.argv[1]
    lib = sys.argv[2] if len(sys.argv) > 2 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                              "libfmftsaxs_b200", "libfmftsaxs.so")
    rows = list(csv.reader(open(sass)))
    hdr, data = rows[1], rows[2:]
    col = {n: hdr.index(n) for n in ("Address", "# Samples", "Instructions Executed", "Thread Instructions Executed",
                                     "stall_long_sb", "stall_barrier", "stall_no_inst", "L2 Theoretical Sectors Local")}
    base = int(data[0][col["Address"]], 16)
    tmp = tempfile.mkdtemp()
    subprocess.run("cd %s && cuobjdump -xelf all %s > /dev/null" % (tmp, lib), shell=True, check=True)
    out = subprocess.run("readelf -sW %s/sxs_exact.sm_100a.cubin 2>/dev/null | grep FUNC" % tmp, shell=True,
                         capture_output=True, text=True).stdout
    subs = []
    for line in out.splitlines():
        p = line.split()
        name = p[-1]
        if "k_fitPK" in name and "$" in name:
            subs.append((int(p[1], 16), int(p[2]), re.sub(r".*\$_Z\d+", "", name)))
    subs.sort()

    def fn(off):
        for v, s, n in subs:
            if v <= off < v + s:
                return re.match(r"[a-z_]+", n).group(0)
        return "k_fit body + objective"

    agg = {}
    for r in data:
        a = agg.setdefault(fn(int(r[col["Address"]], 16) - base), [0] * 7)
        for i, n in enumerate(("# Samples", "Instructions Executed", "Thread Instructions Executed", "stall_long_sb",
                               "stall_barrier", "stall_no_inst", "L2 Theoretical Sectors Local")):
            a[i] += int(r[col[n]] or 0)
    ts = sum(a[0] for a in agg.values())
    ti = sum(a[1] for a in agg.values())
    print("samples %d, warp instructions %d" % (ts, ti))
    print("%-24s %7s %7s %6s %8s %7s %8s %11s" % ("function", "samp%", "inst%", "lanes", "longsb%", "barr%", "noinst%", "locSect(M)"))
    for f, a in sorted(agg.items(), key=lambda z: -z[1][0]):
        print("%-24s %7.1f %7.1f %6.1f %8.1f %7.1f %8.1f %11.1f" % (f, 100 * a[0] / ts, 100 * a[1] / ti, a[2] / max(1, a[1]),
              100 * a[3] / max(1, a[0]), 100 * a[4] / max(1, a[0]), 100 * a[5] / max(1, a[0]), a[6] / 1e6))


if __name__ == "__main__":
    main()
