"""Tuning probe: does running the cross-term kernel of one pose batch concurrently with the fit kernel of another
help?  Two plans on one GPU score the two halves of the z range, first back to back on one stream, then from two
host threads on two streams.  Prints wall ms for both."""
import os
import sys
import threading
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import bench
from libfmftsaxs_b200 import capi


def main():
    nz = int(os.environ.get("PROBE_NZ", "16"))
    nsplit = int(os.environ.get("PROBE_SPLIT", "2"))

    def native_cross(idx, coefA, coefB, q, zv, L):
        pl = capi.Plan(L, q, device=0)
        pl.set_molecules(coefA, coefB)
        pl.set_experiment(np.ones(6 * len(q)), 1.0, 1.0)
        pl.set_translations(zv)
        X = pl.cross_terms(idx)
        pl.close()
        return X

    w = bench.build_inputs(capi.expand, capi.opt_params, native_cross, 0, 70000, nz)
    L, q, zvals, idx = w["L"], w["qvals"], w["zvals"], w["index"]
    n = len(idx)
    dev = torch.device("cuda", 0)
    d_idx = torch.from_numpy(idx).to(dev)
    plans, outs, streams = [], [], []
    for k in range(nsplit):
        p = capi.Plan(L, q, device=0)
        p.set_molecules(w["coefA"], w["coefB"])
        p.set_experiment(w["a"], w["scal"][1], w["scal"][2])
        p.set_translations(zvals)
        plans.append(p)
        outs.append(torch.zeros((3, n), dtype=torch.float64, device=dev))
        streams.append(torch.cuda.Stream())
    bounds = [round(k * nz / nsplit) for k in range(nsplit + 1)]

    def run(k, stream_ptr):
        o = outs[k]
        plans[k].score_device(d_idx.data_ptr(), n, o[0].data_ptr(), o[1].data_ptr(), o[2].data_ptr(), stream_ptr,
                              z_lo=bounds[k], z_hi=bounds[k + 1])

    def serial():
        for k in range(nsplit):
            run(k, 0)
        torch.cuda.synchronize()

    def concurrent():
        th = [threading.Thread(target=run, args=(k, streams[k].cuda_stream)) for k in range(nsplit)]
        for t in th:
            t.start()
        for t in th:
            t.join()
        torch.cuda.synchronize()

    for name, fn in (("serial", serial), ("concurrent", concurrent)):
        for _ in range(2):
            fn()
        t = time.perf_counter()
        for _ in range(3):
            fn()
        print(name, "split", nsplit, "ms", round((time.perf_counter() - t) / 3 * 1e3, 1), flush=True)
    ref = torch.zeros((3, n), dtype=torch.float64, device=dev)
    plans[0].score_device(d_idx.data_ptr(), n, ref[0].data_ptr(), ref[1].data_ptr(), ref[2].data_ptr(), 0)
    torch.cuda.synchronize()
    tot = sum(outs)
    print("max |sum of shards - one call| =", float((tot - ref).abs().max()))


if __name__ == "__main__":
    main()
