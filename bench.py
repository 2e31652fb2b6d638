#!/usr/bin/env python
"""bench.py — conformations scored per second by the FFT-SAXS `correlate` scoring path.

    python bench.py --gpus N --steps K --warmup W            (N>1: launched by torchrun, one rank per GPU)
    python bench.py --impl reference --gpus N --steps K --warmup W

Product arm.  Workload = BASELINE.json configs[2]: synthetic 3000-atom receptor + 1500-atom ligand, L=15, 50 q
points, 70 000 rotations x 64 z steps = 4.48 M poses per GPU (weak scaling: every rank scores its own pose list
of that size, as if the ft files were dealt out; the score tables are all-gathered over NCCL inside the step).
One step = one full pass of sxs_compute_saxs_scores' work over the batch: key sort -> distinct grid points ->
z-translation -> cross terms (K3) -> (c1,c2) fit (K4) -> scatter.
  value : poses/s with the pose list, coefficient tables and outputs resident in HBM (C-ABI *_dev entry),
          timed with CUDA events on the launching stream, max over ranks.
  e2e   : poses/s through the reference-shaped API (sxs_compute_saxs_scores via its flat adapter) with pinned
          HOST buffers in and out — H2D/D2H, the host Bessel table and the host scatter are inside the timing.
Reference arm (--impl reference): the reference's own CPU code (oracle/_ref/libsxsref.so = its unmodified
sources) on the host cores, one process per core over z like its MPI build, on a bounded sample of the same poses.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))

WORKLOAD = "cfg3_3k+1.5k_L15_Q50_70kx64z"
METRIC = "dimer conformations scored/sec (correlate)"
UNIT = "conformations/s"


def env_int(name, dflt):
    try:
        return int(os.environ.get(name, dflt))
    except ValueError:
        return dflt


# ---------------------------------------------------------------------------------------------- clocks

class ClockSampler:
    """nvidia-smi sampled every 200 ms during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-i", str(self.gpu), "-lms", "200"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, smax, reasons = [], [], set()
        for (t, l) in self.lines:
            if t < t0 or t > t1 + 0.3:
                continue
            f = [x.strip() for x in l.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------- workload

def build_inputs(expand, opt_params, native_cross, seed_rank, nrot, nz):
    """molecules identical on every rank, pose list seeded per rank.  The "experimental" curve is the computed
    profile of pose 0 of rank 0's list at c1 = c2 = 1 with 5 % errors — like examples/run_correlate.sh, which
    scores against the single_saxs profile of a reference complex."""
    from libfmftsaxs_b200 import workload as wl
    w = wl.make(WORKLOAD, nrot=nrot, nz=nz)
    if seed_rank:
        w["index"] = wl.make_pose_indices(w["L"], w["zvals"], nrot or wl.CONFIGS[WORKLOAD]["nrot"],
                                          w["seed"] + 3 + 1000 * seed_rank)
    q, L = w["qvals"], w["L"]
    coefA, rmA, _ = expand(wl.MAP_PATH, w["rec"]["xyz"], w["rec"]["res"], w["rec"]["atm"], w["rec"]["radius"], q, L,
                           sa=w["rec"]["sa"], water_mode=1)
    coefB, rmB, _ = expand(wl.MAP_PATH, w["lig"]["xyz"], w["lig"]["res"], w["lig"]["atm"], w["lig"]["radius"], q, L,
                           sa=w["lig"]["sa"], water_mode=1)
    native_idx = wl.make_pose_indices(w["L"], w["zvals"][:1] * 0 + 40.0, 1, w["seed"] + 77)  # one pose at z = 40
    X0 = native_cross(native_idx, coefA, coefB, q, [40.0], L)[0]   # [6][Q]: VV,VD,VW,DD,DW,WW
    I0 = X0[0] - X0[1] + X0[2] + X0[3] - X0[4] + X0[5]              # c1 = c2 = 1 -> G = 1
    eq, ei, ee = q.copy(), I0, 0.05 * I0
    a, scal = opt_params(eq, ei, ee, q, wl.mean_radius(w["rec"], w["lig"]))
    w.update(coefA=coefA, coefB=coefB, a=a, scal=scal)
    return w


def sample_slabs(index, L, nslabs, seed=7):
    """Bounded sample for the CPU arm: all poses of `nslabs` randomly chosen (z, beta2) slabs, i.e. up to L+1
    complete (z, b1, b2) cells each.  The reference's cost is per cell plus one ligand pre-sum per slab, so whole
    slabs keep its amortisation as in a full run.  Returns one row-index array per slab."""
    nb, N = L + 1, 2 * L + 1
    v = index.astype(np.int64) // (N ** 3)          # (z*nb + b1)*nb + b2
    slab = (v // (nb * nb)) * nb + (v % nb)          # z*nb + b2
    if L > 20:
        # one L = 30 cell costs the reference ~25 s: the bounded sample is made of single cells there
        slab = v
    u = np.unique(slab)
    rng = np.random.default_rng(seed)
    pick = rng.choice(u, size=min(nslabs, len(u)), replace=False)
    return [np.flatnonzero(slab == s) for s in pick]


def to_ref_index(index, zvals, L):
    """the reference takes 32-bit flat indices (src/fftsaxs.h:99): trim the z table to the z steps these poses use and
    renumber the z digit, so that any sample of an L = 30 list fits; values of z, and every other digit, stay"""
    nb, N = L + 1, 2 * L + 1
    idx64 = np.asarray(index).astype(np.int64)
    per_z = nb * nb * N ** 3
    zu = np.unique(idx64 // per_z)
    idx = np.searchsorted(zu, idx64 // per_z) * per_z + idx64 % per_z
    if len(idx) and idx.max() >= 2 ** 31:
        raise ValueError("sample spans too many z steps for the reference's 32-bit index")
    return idx.astype(np.int32), np.asarray(zvals)[zu]


# ---------------------------------------------------------------------------------------------- reference arm

def _ref_worker(job):
    import refso
    idx, w = job
    idx, zv = to_ref_index(idx, w["zvals"], w["L"])
    t = time.perf_counter()
    refso.scores(idx, w["coefA"], w["coefB"], w["a"], w["scal"], w["qvals"], zv, w["L"])
    return time.perf_counter() - t


def run_reference(args):
    rank = env_int("RANK", 0)
    if rank != 0:
        return 0
    import refso
    if not refso.available():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libsxsref.so missing (built from /root/reference by oracle/Makefile)"}))
        return 0
    from multiprocessing import Pool
    from golden import proto_cross_terms as proto

    def native_cross(idx, coefA, coefB, q, zv, L):
        A = coefA[..., 0] + 1j * coefA[..., 1]
        B = coefB[..., 0] + 1j * coefB[..., 1]
        return proto.cross_terms(idx, A, B, q, zv, L, proto.reference_tables(L, q, zv))

    w = build_inputs(refso.expand, refso.opt_params, native_cross, 0, args.nrot, args.nz)
    L = w["L"]
    N = 2 * L + 1
    cores = os.cpu_count() or 1
    procs = max(1, min(cores, args.ref_procs or cores))
    # bounded sample: one (z, beta2) slab per process — the processes work on different z like the MPI ranks of
    # tools/correlate.c:142-147, each paying the full per-call set-up (tables, allocations) like a rank does
    shards = [w["index"][r] for r in sample_slabs(w["index"], L, procs * args.ref_slabs_per_proc)]
    shards = [np.concatenate(shards[i::procs]) for i in range(min(procs, len(shards)))]
    idx = np.concatenate(shards)
    small = {k: w[k] for k in ("coefA", "coefB", "a", "scal", "qvals", "zvals", "L")}
    times = []
    with Pool(len(shards)) as pool:
        for it in range(args.warmup + args.steps):
            t = time.perf_counter()
            pool.map(_ref_worker, [(s, small) for s in shards])
            dt = time.perf_counter() - t
            if it >= args.warmup:
                times.append(dt)
    tot = sum(times)
    value = len(idx) * len(times) / tot
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * tot / len(times), "higher_is_better": True,
            # the CPU figure does not depend on the number of GPUs; the label follows the product arm's mode for this N
            "scaling": (args.scaling if args.gpus > 1 else "weak"),
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "poses_in_sample": int(len(idx)), "cells_in_sample": int(len(np.unique(idx.astype(np.int64) // N ** 3))),
                       "L": L, "qnum": len(w["qvals"])},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": len(shards), "kind": "reference",
                             "sample": "all poses of %d random (z,beta2) slabs (each up to L+1 complete cells) of the workload, one "
                                       "slab per process, %d processes; reference sources unmodified, FFTW replaced by a "
                                       "split-radix-free DFT shim (oracle/shim/fftw_shim.c)" % (procs * args.ref_slabs_per_proc, len(shards))},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    return 0


# ---------------------------------------------------------------------------------------------- product arm

def run_product(args):
    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    os.environ["SXS_CUDA_DEVICES"] = str(local)
    import torch
    import torch.distributed as dist
    from libfmftsaxs_b200 import capi

    if not torch.cuda.is_available() or capi.device_count() < 1:
        raise SystemExit("bench.py: no CUDA device — the product has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    cpu_group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        cpu_group = dist.new_group(backend="gloo")   # host-side barrier for the e2e leg (no spinning kernels on the GPUs)

    def native_cross(idx, coefA, coefB, q, zv, L):
        pl = capi.Plan(L, q, device=local)
        pl.set_molecules(coefA, coefB)
        pl.set_experiment(np.ones(6 * len(q)), 1.0, 1.0)
        pl.set_translations(zv)
        X = pl.cross_terms(idx)
        pl.close()
        return X

    strong = args.scaling == "strong" and world > 1
    w = build_inputs(capi.expand, capi.opt_params, native_cross, 0 if strong else rank, args.nrot, args.nz)
    L, q, zvals, idx = w["L"], w["qvals"], w["zvals"], w["index"]
    full_idx = idx
    gsh = None
    if strong:
        # ONE list over the ranks: every rank filters the rows of its own z range (the reference's MPI scheme,
        # tools/correlate.c:140-147) and the score table is gathered into input order inside the timed step
        from libfmftsaxs_b200 import dist as sd
        gsh = sd.RowShardGather(full_idx, L, len(zvals), world, rank, device=dev)
        idx = np.ascontiguousarray(full_idx[gsh.my_rows])
    n = len(idx)
    n_job = len(full_idx) if strong else world * n   # poses the whole job scores per step
    plan = capi.Plan(L, q, device=local)
    plan.set_molecules(w["coefA"], w["coefB"])
    plan.set_experiment(w["a"], w["scal"][1], w["scal"][2])
    plan.set_translations(zvals)

    is64 = idx.dtype == np.int64
    d_idx = torch.from_numpy(idx).to(dev)
    d_out = gsh.local if strong else torch.zeros((3, n), dtype=torch.float64, device=dev)
    gathered = torch.empty((world * 3, n), dtype=torch.float64, device=dev) if (world > 1 and not strong) else None
    stream = torch.cuda.current_stream().cuda_stream

    def step():
        plan.score_device(d_idx.data_ptr(), n, d_out[0].data_ptr(), d_out[1].data_ptr(), d_out[2].data_ptr(), stream, i64=is64)
        if strong:
            gsh.gather()
        elif world > 1:
            dist.all_gather_into_tensor(gathered, d_out)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    plan.set_profiling(True)
    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.3)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t0 = time.time()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    t1 = time.time()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop(t0, t1)
    ktimes = plan.kernel_times()
    plan.set_profiling(False)
    stats = plan.stats()
    hist = plan.fit_evaluations()
    nfg_mean = float((hist * np.arange(64)).sum() / max(1, hist.sum()))
    if world > 1:
        tt = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms = float(tt.item())
    value = n_job * args.steps / (ms * 1e-3)

    # ---- end to end through the reference-shaped API with host buffers ----
    # strong scaling: rank 0 alone calls the API on the WHOLE list with every GPU of the job named in SXS_CUDA_DEVICES
    # (one host thread per device inside the library: the drop-in replacement of `mpirun -np N correlate`); the other
    # ranks wait at the barrier.  Weak scaling: every rank calls the API on its own list.
    e2e_idx = full_idx if strong else idx
    e2e_n = len(e2e_idx)
    if strong:
        os.environ["SXS_CUDA_DEVICES"] = ",".join(str(d) for d in range(world))
    h_idx = torch.from_numpy(e2e_idx).pin_memory()
    h_out = [torch.zeros(e2e_n, dtype=torch.float64).pin_memory() for _ in range(3)]
    import ctypes as C
    lib = capi.lib()
    dp = C.POINTER(C.c_double)
    cA, cB, a_, sc_, q_, z_ = [np.ascontiguousarray(x, dtype=np.float64) for x in (w["coefA"], w["coefB"], w["a"], w["scal"], q, zvals)]

    def e2e_step():
        args_tail = (capi.dptr(cA), capi.dptr(cB), capi.dptr(a_), capi.dptr(sc_), capi.dptr(q_), C.c_int(len(q_)),
                     capi.dptr(z_), C.c_int(len(z_)), C.c_int(L), C.c_int(1))
        o = [C.cast(t.data_ptr(), dp) for t in h_out]
        if strong and rank != 0:
            return 0.0
        if is64:
            lib.sxs_flat_scores64(o[0], o[1], o[2], C.cast(h_idx.data_ptr(), C.POINTER(C.c_longlong)), C.c_longlong(e2e_n), *args_tail)
        else:
            lib.sxs_flat_scores(o[0], o[1], o[2], C.cast(h_idx.data_ptr(), C.POINTER(C.c_int)), C.c_int(e2e_n), *args_tail)
        return float(h_out[0][0])

    def host_barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier(group=cpu_group)

    e2e_steps = max(1, min(args.steps, 3))
    e2e_step()
    host_barrier()
    te = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    host_barrier()
    e2e_s = time.perf_counter() - te
    if world > 1:
        tt = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_s = float(tt.item())
    e2e_value = n_job * e2e_steps / e2e_s
    N = 2 * L + 1
    tables = cA.nbytes + cB.nbytes + a_.nbytes + len(z_) * len(q_) * N * 8
    if strong:
        h2d = full_idx.nbytes + world * tables      # every device gets the tables, each its own rows
        d2h = len(full_idx) * (3 * 8 + 4)
    else:
        h2d = idx.nbytes + tables
        d2h = n * (3 * 8 + 4)

    # parity spot check of what was just timed: resident path == host path (same kernels)
    if strong:
        same = bool(rank != 0 or np.array_equal(gsh.table[0, :len(full_idx)].cpu().numpy(), h_out[0].numpy()))
    else:
        same = bool(np.array_equal(d_out[0].cpu().numpy(), h_out[0].numpy()))

    line = None
    if rank == 0:
        ML = (L + 1) * (L + 2) // 2
        Q = len(q)
        fp64_peak = capi.fp64_peak(local)
        peak_src = "DFMA microbenchmark run by this bench on this GPU (MEASURED_PEAKS.json has no FP64 entry)"
        # K3: 9*ML complex MACs (8 flops) + (L+1) phase steps of 6 real MAC pairs, per pose and q
        flops_per_point = Q * (9 * ML * 8 + (L + 1) * 6 * 4)
        cross_ms, cross_n = ktimes["cross"]
        fit_ms, fit_n = ktimes["fit"]
        evals_total = float((hist * np.arange(64)).sum())
        # K4: per objective evaluation and q node 147 flops + 2 exp at 20 flops, optimiser logic not counted
        fit_flops = evals_total * Q * (147 + 2 * 20)

        traffic = ncu_traffic(n)

        def roof(name, flops_per_step, ms_total, launches):
            ach = flops_per_step * args.steps / (ms_total * 1e-3) / 1e12 if ms_total > 0 else None
            key = name.split(" ")[0]
            dram = traffic.get(key)
            return {"kernel": name, "bound": "fp64", "achieved": ach, "peak": fp64_peak, "unit": "TFLOP/s",
                    "frac": (ach / fp64_peak) if ach else None, "traffic": dram, "peak_source": peak_src,
                    # measured DRAM bytes per launch over the live launch time, beside the copy bandwidth of MEASURED_PEAKS.json
                    "dram_gbs": (dram / (ms_total / max(1, launches) * 1e-3) / 1e9) if (dram and ms_total > 0) else None,
                    "hbm_peak_gbs": hbm_peak(),
                    "traffic_source": TRAFFIC_FILE if key in traffic else None,
                    "flops_per_launch": flops_per_step / max(1, launches // args.steps),
                    "avg_launch_ms": ms_total / max(1, launches), "ms_per_step": ms_total / args.steps}

        r_cross = roof("k_cross (K3: angular transform + cross terms)", stats["points"] * flops_per_point, cross_ms, cross_n)
        r_cross["l1_side_gbs"] = stats["points"] * args.steps * (2 * 3 * ML * 16 * Q) / (cross_ms * 1e-3) / 1e9 if cross_ms > 0 else None
        r_fit = roof("k_fit (K4: fused (c1,c2) fit)", fit_flops, fit_ms, fit_n)
        r_fit["evaluations_per_fit"] = nfg_mean
        roofline, roofline2 = (r_fit, r_cross) if fit_ms >= cross_ms else (r_cross, r_fit)
        cpu, parity_line = cpu_baseline(w, args, d_out)
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic",
                "config": {"workload": WORKLOAD, "poses_per_gpu": int(n), "poses_per_step_whole_job": int(n_job),
                           "list": ("one list of %d poses split over %d GPUs by z range, score table gathered into input order "
                                    "inside the step" % (len(full_idx), world)) if strong else "one list per GPU",
                           "distinct_grid_points": int(stats["points"]),
                           "L": L, "qnum": Q, "z_steps": int(len(zvals)), "rec_atoms": len(w["rec"]["res"]),
                           "lig_atoms": len(w["lig"]["res"]),
                           "l2_policy": "inputs larger than L2 (rotated tables %.0f MB + translated slabs %.1f GB per step)"
                                        % (2 * (L + 1) * Q * 3 * ML * N * 16 / 1e6, stats["slabs"] * Q * 3 * ML * N * 16 / 1e9)},
                "clocks": clocks,
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                        "steps": e2e_steps, "api": "sxs_compute_saxs_scores (flat adapter), pinned host buffers"
                        + (", ONE call on rank 0 driving all %d GPUs (SXS_CUDA_DEVICES)" % world if strong else "")},
                "gpu_launches": int(stats["launches"] * args.steps),
                "kernels_ms_per_step": {k: v[0] / args.steps for k, v in ktimes.items()},
                "fit_evaluations": {"mean": nfg_mean, "p50": int(np.searchsorted(np.cumsum(hist), 0.5 * hist.sum())),
                                    "p99": int(np.searchsorted(np.cumsum(hist), 0.99 * hist.sum())),
                                    "max_bin": int(np.flatnonzero(hist).max()) if hist.sum() else 0},
                "pose_list": pose_list_stats(idx, L),
                "roofline": roofline, "roofline_second_kernel": roofline2, "cpu_baseline": cpu, "parity": parity_line,
                "value_excludes": "plan set-up (rotated coefficient tables: k_rotate x2, 0.8 ms; L-only tables, cached per "
                                  "process; j_p(qz) table) happens before the timed steps; the e2e leg calls "
                                  "sxs_compute_saxs_scores, which includes it on every call",
                "resident_equals_host_path": same}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def pose_list_stats(idx, L):
    """the distributions throughput depends on: rows per (z, beta1, beta2) cell"""
    N = 2 * L + 1
    _, counts = np.unique(idx.astype(np.int64) // N ** 3, return_counts=True)
    return {"cells": int(len(counts)), "rows_per_cell_mean": float(counts.mean()),
            "rows_per_cell_p50": int(np.percentile(counts, 50)), "rows_per_cell_p99": int(np.percentile(counts, 99)),
            "rows_per_cell_max": int(counts.max())}


TRAFFIC_FILE = "profiles/r2al_ncu_traffic.json"


def hbm_peak():
    try:
        return json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json"))).get("hbm_gbs")
    except (OSError, ValueError):
        return None


def ncu_traffic(poses_per_gpu):
    """DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of the kernels, from the committed ncu
    launch list of this very command; only quoted when the workload is the one that was profiled"""
    try:
        t = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), TRAFFIC_FILE)))
    except OSError:
        return {}
    if t.get("workload") != WORKLOAD or t.get("poses_per_gpu") != poses_per_gpu:
        return {}
    return {k: v["dram_read_bytes_per_launch"] + v["dram_write_bytes_per_launch"] for k, v in t["kernels"].items()}


def _sens_worker(job):
    import refso
    idx, w = job
    return refso.scores(idx, w["coefA"], w["coefB"], w["a"], w["scal"], w["qvals"], w["zvals"], w["L"], so=refso.SENS_SO)


def cpu_baseline(w, args, d_out):
    """The compiled reference on ONE host core over a bounded sample of the same poses (rank 0, N=1 only) — and what
    the reference says about those poses is compared with what the GPU just wrote for them (`parity`).  Beside it,
    `reference_self_noise`: the same sample scored by the reference's own sources built with FMA contraction
    (oracle/Makefile, libsxsref_fma.so), i.e. how far the reference is from itself under a 1e-16 perturbation."""
    if env_int("WORLD_SIZE", 1) != 1 or args.no_cpu_baseline:
        return None, None
    import refso
    import parity
    if not refso.available():
        return {"value": None, "unit": UNIT, "cores": 0, "kind": "reference", "sample": "oracle/_ref missing"}, None
    L = w["L"]
    rows = np.concatenate(sample_slabs(w["index"], L, args.cpu_slabs))
    idx, zv = to_ref_index(w["index"][rows], w["zvals"], L)
    small = dict(coefA=w["coefA"], coefB=w["coefB"], a=w["a"], scal=w["scal"], qvals=w["qvals"], zvals=zv, L=L)
    pool = None
    pending = None
    if refso.sens_available() and not args.no_self_noise:
        from multiprocessing import get_context
        pool = get_context("spawn").Pool(1)
        pending = pool.apply_async(_sens_worker, ((idx, small),))
    t = time.perf_counter()
    want = refso.scores(idx, small["coefA"], small["coefB"], small["a"], small["scal"], small["qvals"], small["zvals"], L)
    dt = time.perf_counter() - t
    import torch
    got = d_out[:, torch.from_numpy(rows).to(d_out.device)].cpu().numpy()
    par = parity.summary((got[0], got[1], got[2]), want)
    par["against"] = "compiled reference (oracle/_ref/libsxsref.so) on the cpu_baseline sample of the benchmarked list"
    if pending is not None:
        sens = pending.get()
        pool.close()
        noise = parity.summary(sens, want)
        par["reference_self_noise"] = {k: noise[k] for k in ("max_rel_chi", "max_rel_c1", "max_rel_c2", "rows_over_tol")}
    cpu = {"value": len(idx) / dt, "unit": UNIT, "cores": 1, "kind": "reference", "seconds": dt,
           "sample": "all %d poses of %d random (z,beta2) slab(s) of the workload (up to L+1 complete cells each); "
                     "unmodified reference sources, FFTW replaced by the DFT shim" % (len(idx), args.cpu_slabs)}
    return cpu, par


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=None, help="pose-list/molecule configuration of libfmftsaxs_b200.workload.CONFIGS "
                    "(default: BASELINE config 3; config 4 is a separate measurement, see profiles/)")
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"],
                    help="N > 1: strong = ONE pose list of the workload's size split over the GPUs (default; what correlate "
                         "under MPI does); weak = every GPU scores its own list of that size")
    ap.add_argument("--nrot", type=int, default=None, help="override rotations per z (default 70000)")
    ap.add_argument("--nz", type=int, default=None, help="override number of z steps (default 64)")
    ap.add_argument("--cpu-slabs", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-self-noise", action="store_true", help="skip the FMA-build leg of the parity report")
    ap.add_argument("--ref-procs", type=int, default=0)
    ap.add_argument("--ref-slabs-per-proc", type=int, default=1)
    args = ap.parse_args()
    if args.workload:
        global WORKLOAD
        WORKLOAD = args.workload
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        return run_reference(args)
    return run_product(args)


if __name__ == "__main__":
    sys.exit(main())
