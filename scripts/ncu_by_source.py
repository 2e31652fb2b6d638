"""Joins the SASS page of an ncu report with nvdisasm's line info of the same build and prints, per source file and per
function range, the executed warp instructions, active lanes and stall samples of one kernel.

    python scripts/ncu_by_source.py REPORT.ncu-rep OBJECT.o KERNEL_SUBSTRING [top_n_lines]
"""
import csv
import io
import os
import re
import subprocess
import sys
import tempfile


def main():
    rep, obj, kern = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 30
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = None
    sass = []
    for r in rows:
        if r and r[0] == "Address":
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            sass.append(dict(zip(hdr, r)))
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
    cub = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cub)], capture_output=True, text=True).stdout.split("\n")
    inside = False
    cur = ("?", 0)
    lines = []
    for l in dis:
        if l.startswith(".text."):
            inside = kern in l
            continue
        if not inside:
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        if re.match(r"\s*/\*[0-9a-f]{4,}\*/", l):
            lines.append(cur)
    n = min(len(lines), len(sass))
    print("sass instructions: report %d, disassembly %d" % (len(sass), len(lines)))
    byfile, byline = {}, {}
    tot_i = tot_s = 0
    for k in range(n):
        d = sass[k]
        ie = float(d["Instructions Executed"] or 0)
        te = float(d["Thread Instructions Executed"] or 0)
        sm = float(d["# Samples"] or 0)
        lsb = float(d.get("stall_long_sb") or 0)
        tot_i += ie
        tot_s += sm
        for key, dct in ((lines[k][0], byfile), (lines[k], byline)):
            a = dct.setdefault(key, [0, 0, 0, 0])
            a[0] += ie; a[1] += te; a[2] += sm; a[3] += lsb
    print("total warp instructions %.4e, samples %d" % (tot_i, tot_s))
    for key, a in sorted(byfile.items(), key=lambda x: -x[1][0]):
        print("%-20s instr %.3e (%4.1f %%)  lanes %4.1f  samples %4.1f %%  long_sb %4.1f %% of its samples"
              % (key, a[0], 100 * a[0] / tot_i, a[1] / max(a[0], 1), 100 * a[2] / max(tot_s, 1), 100 * a[3] / max(a[2], 1)))
    print("-- top lines by samples")
    for key, a in sorted(byline.items(), key=lambda x: -x[1][2])[:top]:
        print("%5.1f %% samples  %5.2f %% instr  lanes %4.1f  long_sb %4.1f %%  %s:%d"
              % (100 * a[2] / tot_s, 100 * a[0] / tot_i, a[1] / max(a[0], 1), 100 * a[3] / max(a[2], 1), key[0], key[1]))


if __name__ == "__main__":
    main()
