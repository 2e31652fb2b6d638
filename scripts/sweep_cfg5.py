#!/usr/bin/env python
"""BASELINE config 5: throughput against lmax and pose count (SURVEY.md §8d "Config 5"), on one GPU or — launched by
torchrun, one rank per GPU — with every list split over the GPUs by z range and its score table gathered into input
order inside the timed call (strong scaling, libfmftsaxs_b200/dist.py:RowShardGather).

    python scripts/sweep_cfg5.py --L 10,15,20,30,40 --rows 1e5,1e6,1e7 [--out gpurun_out/cfg5.jsonl]
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 scripts/sweep_cfg5.py ...

Molecules and curve as config 3 (3000 + 1500 synthetic atoms, Q = 50, 64 z steps 17..80); rows = rotations x 64.
Per point: poses/s with everything resident (CUDA events around the score call, 1 warm-up + 2 timed calls), per-kernel
device times, K3/K4 share of the measured FP64 peak, rows per cell, objective evaluations per fit.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--L", default="10,15,20,30,40")
    ap.add_argument("--rows", default="1e5,1e6,1e7")
    ap.add_argument("--out", default=os.path.join(REPO, "gpurun_out", "cfg5.jsonl"))
    ap.add_argument("--steps", type=int, default=2)
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    os.environ["SXS_CUDA_DEVICES"] = str(local)
    import torch
    import torch.distributed as dist
    from libfmftsaxs_b200 import capi
    from libfmftsaxs_b200 import workload as wl
    from libfmftsaxs_b200 import dist as sd

    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    peak = capi.fp64_peak(local)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    fout = open(args.out, "a") if rank == 0 else open(os.devnull, "w")
    base = wl.make("cfg3_3k+1.5k_L15_Q50_70kx64z", nrot=1, nz=1)
    rec, lig = base["rec"], base["lig"]
    zvals = wl.CONFIGS["cfg3_3k+1.5k_L15_Q50_70kx64z"]["zvals"]
    q = base["qvals"]
    Q = len(q)
    for L in [int(x) for x in args.L.split(",")]:
        t0 = time.time()
        A, _, _ = capi.expand(wl.MAP_PATH, rec["xyz"], rec["res"], rec["atm"], rec["radius"], q, L, sa=rec["sa"], water_mode=1)
        B, _, _ = capi.expand(wl.MAP_PATH, lig["xyz"], lig["res"], lig["atm"], lig["radius"], q, L, sa=lig["sa"], water_mode=1)
        plan = capi.Plan(L, q, device=local)
        plan.set_molecules(A, B)
        one = wl.make_pose_indices(L, np.array([40.0]), 1, base["seed"] + 77)
        plan.set_experiment(np.ones(6 * Q), 1.0, 1.0)
        plan.set_translations(np.array([40.0]))
        X0 = plan.cross_terms(one)[0]
        I0 = X0[0] - X0[1] + X0[2] + X0[3] - X0[4] + X0[5]
        a, scal = capi.opt_params(q.copy(), I0, 0.05 * I0, q, wl.mean_radius(rec, lig))
        plan.set_experiment(a, scal[1], scal[2])
        plan.set_translations(zvals)
        setup_s = time.time() - t0
        ML = (L + 1) * (L + 2) // 2
        for rows in [int(float(x)) for x in args.rows.split(",")]:
            nrot = max(1, rows // len(zvals))
            t1 = time.time()
            full = wl.make_pose_indices(L, zvals, nrot, base["seed"] + 3)
            gen_s = time.time() - t1
            n_total = len(full)
            gsh = sd.RowShardGather(full, L, len(zvals), world, rank, device=dev)
            idx = np.ascontiguousarray(full[gsh.my_rows])
            n = len(idx)
            is64 = idx.dtype == np.int64
            d_idx = torch.from_numpy(idx).to(dev)
            d_out = gsh.local
            stream = torch.cuda.current_stream().cuda_stream

            def step():
                if n:
                    plan.score_device(d_idx.data_ptr(), n, d_out[0].data_ptr(), d_out[1].data_ptr(), d_out[2].data_ptr(), stream, i64=is64)
                gsh.gather()

            def barrier():
                torch.cuda.synchronize()
                if world > 1:
                    dist.barrier()
                torch.cuda.synchronize()

            step()
            barrier()
            plan.set_profiling(True)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.steps):
                step()
            e1.record()
            barrier()
            ms = e0.elapsed_time(e1) / args.steps
            if world > 1:
                tt = torch.tensor([ms], dtype=torch.float64, device=dev)
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
                ms = float(tt.item())
            kt = plan.kernel_times()
            plan.set_profiling(False)
            st = plan.stats()
            hist = plan.fit_evaluations()
            evals = float((hist * np.arange(64)).sum())
            cross_ms, fit_ms = kt["cross"][0] / args.steps, kt["fit"][0] / args.steps
            k3_flops = st["points"] * Q * (9 * ML * 8 + (L + 1) * 6 * 4)
            k4_flops = evals * Q * (147 + 2 * 20)
            cells = len(np.unique(full.astype(np.int64) // (2 * L + 1) ** 3))
            ok = bool(torch.isfinite(gsh.table[:, :n_total]).all().item())
            line = {"L": L, "rows": n_total, "n_gpus": world, "rows_rank0": n, "poses_per_s": n_total / (ms * 1e-3), "ms_per_call": ms,
                    "distinct_points": int(st["points"]),
                    "cells": int(cells), "rows_per_cell": n_total / cells, "z_groups": int(st["groups"]), "index_bits": 64 if is64 else 32,
                    "kernels_ms": {k: v[0] / args.steps for k, v in kt.items()},
                    "k3_tflops": k3_flops / (cross_ms * 1e-3) / 1e12 if cross_ms > 0 else None,
                    "k3_frac_fp64_peak": k3_flops / (cross_ms * 1e-3) / 1e12 / peak if cross_ms > 0 else None,
                    "k4_tflops": k4_flops / (fit_ms * 1e-3) / 1e12 if fit_ms > 0 else None,
                    "k4_frac_fp64_peak": k4_flops / (fit_ms * 1e-3) / 1e12 / peak if fit_ms > 0 else None,
                    "evaluations_per_fit": evals / max(1, st["points"]), "fp64_peak_tflops": peak, "finite": ok,
                    "setup_s": setup_s, "index_generation_s": gen_s}
            if rank == 0:
                print(json.dumps(line), flush=True)
            fout.write(json.dumps(line) + "\n")
            fout.flush()
            del d_idx, d_out, gsh
            torch.cuda.empty_cache()
        plan.close()
    fout.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
