#!/usr/bin/env python
"""Per-kernel summary of an ncu launch list (--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
--csv): launches, mean duration, DRAM bytes read/written per launch, share of the listed GPU time.

    python scripts/ncu_traffic.py gpurun_out/launches_r1e.csv profiles/r1e_ncu_traffic.json [poses_per_gpu]

Launches shorter than half of a kernel's longest one are left out of its averages (the bench runs every kernel once on
a single pose while it builds its synthetic experimental curve)."""
import csv
import json
import sys
from collections import OrderedDict, defaultdict


def main():
    src, dst = sys.argv[1], sys.argv[2]
    poses = int(sys.argv[3]) if len(sys.argv) > 3 else 4480000
    rows = [r for r in csv.reader(l for l in open(src) if l.startswith('"'))]
    hdr, rows = rows[0], rows[1:]
    col = {n: i for i, n in enumerate(hdr)}
    per = OrderedDict()
    for r in rows:
        key = r[col["ID"]]
        per.setdefault(key, {"name": r[col["Kernel Name"]].split("(")[0].split("<")[0].replace("void ", "").strip()})
        v = float(r[col["Metric Value"]].replace(",", ""))
        unit = r[col["Metric Unit"]]
        m = r[col["Metric Name"]]
        if m == "gpu__time_duration.sum":
            v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
        elif unit in ("Kbyte", "Mbyte", "Gbyte", "Tbyte"):
            v *= {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[unit]
        per[key][m] = v
    by = defaultdict(list)
    for k in per.values():
        by[k["name"]].append(k)
    total_ms = sum(k.get("gpu__time_duration.sum", 0.0) for k in per.values())
    out = {"source": src, "workload": "cfg3_3k+1.5k_L15_Q50_70kx64z", "poses_per_gpu": poses, "listed_gpu_ms": total_ms, "kernels": {}}
    for name, ls in sorted(by.items(), key=lambda kv: -sum(x.get("gpu__time_duration.sum", 0.0) for x in kv[1])):
        # full-size launches only: the bench also runs each kernel once on a single pose to make its synthetic curve
        top = max(x.get("gpu__time_duration.sum", 0.0) for x in ls)
        ls = [x for x in ls if x.get("gpu__time_duration.sum", 0.0) >= 0.5 * top]
        n = len(ls)
        ms = sum(x.get("gpu__time_duration.sum", 0.0) for x in ls)
        out["kernels"][name] = {
            "launches": n,
            "ncu_ms_per_launch": ms / n,
            "share_of_listed_gpu_time": ms / total_ms if total_ms else None,
            "dram_read_bytes_per_launch": sum(x.get("dram__bytes_read.sum", 0.0) for x in ls) / n,
            "dram_write_bytes_per_launch": sum(x.get("dram__bytes_write.sum", 0.0) for x in ls) / n,
        }
    json.dump(out, open(dst, "w"), indent=1)
    for name, k in list(out["kernels"].items())[:8]:
        print("%-22s n=%3d  %8.3f ms  %5.1f %%  rd %7.2f GB  wr %7.2f GB" % (name, k["launches"], k["ncu_ms_per_launch"],
              100 * k["share_of_listed_gpu_time"], k["dram_read_bytes_per_launch"] / 1e9, k["dram_write_bytes_per_launch"] / 1e9))


if __name__ == "__main__":
    main()
