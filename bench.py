#!/usr/bin/env python
"""bench.py -- atom-timesteps/s of the ReaxFF+QEq hot path on N B200s (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this framework (one rank per GPU)
    python bench.py --impl reference --steps K --warmup W    # CPU restatement of the reference on the host cores

A "step" is one pass of the reference's main-loop body (src/main.F90:64-98): integrator halves, COPYATOMS(MOVE),
QEq (CG to QEq_tol 1e-7) and FORCE over one synthetic RDX configuration.  N=1 workload = BASELINE.json configs[1]:
conf/init.rdx.lg (LG force field) replicated 18x18x18 = 979 776 atoms, Gaussian sigma=0.02 A displacements
(seed 20261017), zero initial velocities and charges.  N>1: weak scaling, the same 18^3 block per GPU, vprocs
(2,1,1) (2,2,1) (2,2,2); `--strong` keeps the 18^3 block in total and splits it over the ranks instead.

value : device-resident stepping (rxg_md_run), inputs in HBM when the clock starts, timed with CUDA events on the
        library's own stream, max over ranks.
e2e   : the same step driven through the reference-facing entry points COPYATOMS(MODE_MOVE) / QEq / FORCE with
        pinned HOST arrays (host<->device copies inside the timed region, integrator on the host like the Fortran
        driver).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

UTIME = 1.0e3 / 20.455            # reference src/module.F90:202
DT_FS = 0.25                       # README sample run
LEX_K = 2.0
INPUTS = os.path.join(ROOT, "tests", "golden", "inputs")
VPROCS = {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}


def workload(args, nranks):
    from rxmd_b200.host.system import build_system
    vp = VPROCS[nranks]
    # weak scaling (default): --mc unit cells per GPU; --strong: --mc unit cells in total, split over the ranks
    mc = tuple(args.mc) if args.strong else tuple(args.mc[a] * vp[a] for a in range(3))
    g = os.path.join(INPUTS, "init.rdx.lg")
    s = build_system(os.path.join(g, "input.xyz"), os.path.join(g, "ffield"), mc=mc, vprocs=vp, isLG=True,
                     displace_sigma=args.sigma)
    return s, mc, vp


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False
        self.proc = None

    def run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, text=True)
            for line in self.proc.stdout:
                self.rows.append([x.strip() for x in line.split(",")])
                if self.stop_flag:
                    break
        except Exception:
            pass

    def finish(self):
        self.stop_flag = True
        if self.proc:
            self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            for k, nm in enumerate(names):
                if len(r) > 3 + k and r[3 + k].lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def peak_hbm():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def cpu_sample(args, steps, warmup):
    """The CPU restatement of the reference (oracle/, OpenMP) on a bounded sample of the same workload."""
    from rxmd_b200.host.system import build_system
    from oracle.pyoracle import Oracle
    g = os.path.join(INPUTS, "init.rdx.lg")
    mc = tuple(args.cpu_mc)
    s = build_system(os.path.join(g, "input.xyz"), os.path.join(g, "ffield"), mc=mc, isLG=True, displace_sigma=args.sigma)
    o = Oracle(s, s.config())
    dt = DT_FS / UTIME
    lw2 = 2.0 * LEX_K / dt / dt
    o.move(); o.qeq(); o.force()
    o.md_run(warmup, dt, 1, lw2, 0)
    t0 = time.perf_counter()
    o.md_run(steps, dt, 1, lw2, warmup)
    t = time.perf_counter() - t0
    cores = int(os.environ.get("OMP_NUM_THREADS", os.cpu_count() or 1))
    o.close()
    return {"value": s.natoms * steps / t, "unit": "atom-timesteps/s", "cores": cores, "kind": "port",
            "sample": f"RDX (LG ffield) {mc[0]}x{mc[1]}x{mc[2]} = {s.natoms} atoms, {steps} steps after {warmup} warm-up, "
                      f"sigma={args.sigma} A; OpenMP C++ restatement of the reference loops (no Fortran toolchain in the image)"}, t / steps * 1e3


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, min(args.steps, 20))
    warm = max(1, min(args.warmup, 3))
    cb, ms = cpu_sample(args, steps, warm)
    line = {"impl": "reference", "metric": "atom-timesteps/s, RDX ReaxFF+QEq", "value": cb["value"], "unit": "atom-timesteps/s",
            "n_gpus": args.gpus, "steps": steps, "warmup": warm, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "RDX conf/init.rdx.lg (LG ffield) ReaxFF+QEq, QEq_tol 1e-7 every step, NVE dt 0.25 fs", "sample": cb["sample"]},
            "cpu_baseline": cb, "e2e": {"value": cb["value"], "unit": "atom-timesteps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--mc", type=int, nargs=3, default=[18, 18, 18], help="unit-cell replication per GPU")
    ap.add_argument("--cpu-mc", type=int, nargs=3, default=[6, 6, 6])
    ap.add_argument("--sigma", type=float, default=0.02)
    ap.add_argument("--strong", action="store_true", help="strong scaling: --mc is the TOTAL replication, split over the GPUs")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run --nproc-per-node N for --gpus N > 1")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    from rxmd_b200.host.engine import Engine, MODE_MOVE
    # e2e leg: let rxg_force reuse the halo and 10 A list of the rxg_qeq that precedes it when the host hands back
    # bit-identical atoms (verified on the device); rxg_md_run does the same sharing internally
    os.environ.setdefault("RXG_FUSE_API", "1")
    s, mc, vp = workload(args, world)
    cfg = s.config(device=local)
    e = Engine(s, cfg, rank=rank)
    if world > 1:
        e.comm_init_torch(dist)
    st = s.ranks[rank]
    nres = len(st["atype"])
    dt = DT_FS / UTIME
    lw2 = 2.0 * LEX_K / dt / dt

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def allmax(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def allsum(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    # ---------------- device-resident stepping: `value`
    atype, pos, v, f, q = e.host_arrays(st)
    e.state_upload(atype, pos, v, q)
    e.md_prime()
    e.md_run(args.warmup, dt, 1, lw2, 0)
    t_before = e.timers()
    l_before = e.launches()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    e.md_run(args.steps, dt, 1, lw2, args.warmup)
    barrier()
    clocks = sampler.finish() if rank == 0 else None
    t_after = e.timers()
    l_after = e.launches()
    ms_total = allmax(t_after[3] - t_before[3])
    natoms_total = allsum(float(e.natoms_resident()))
    value = natoms_total * args.steps / (ms_total * 1e-3)
    d = t_after - t_before
    cg_iters = d[17] / max(args.steps, 1)
    pe, ke, qsum, _ = e.md_observe()

    # ---------------- roofline of the dominant kernel: the get_hsh SpMV (+ the get_gradient SpMV beside it)
    nnz, nloc, ntot = t_after[14], t_after[15], t_after[16]
    peak, peak_src = peak_hbm()

    def roof(ms_sum, launches, k):
        if launches <= 0:
            return None
        bytes_alg = 12.0 * nnz + 4.0 * (nloc + 1) + 8.0 * k * ntot + 40.0 * nloc
        ach = bytes_alg / (ms_sum / launches * 1e-3) / 1e9
        return bytes_alg, ach
    traffic = None
    tp = os.path.join(ROOT, "profiles", "spmv_traffic.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    r_h = roof(d[10], d[11], 2)      # the single-pass CG gathers two vectors (hs, ht)
    r_g = roof(d[12], d[13], 2)
    roofline = None
    if r_h:
        kname = ("k_spmv_rows16 (QEq CG SpMV H.(hs,ht): TMA-staged fp64 values + 16-bit column stream, 16 lanes per row)" if t_after[19] > 0
                 else "k_spmv_rows (QEq CG SpMV H.(hs,ht): TMA-staged matrix stream, 16 lanes per row)")
        roofline = {"kernel": kname, "bound": "hbm", "achieved": r_h[1], "peak": peak,
                    "unit": "GB/s", "frac": r_h[1] / peak, "traffic": traffic, "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": r_h[0], "avg_launch_ms": d[10] / d[11], "launches_timed": int(d[11]),
                    "step_share": d[10] / max(t_after[3] - t_before[3], 1e-9),
                    "k_gradient": {"achieved": r_g[1], "frac": r_g[1] / peak, "avg_launch_ms": d[12] / d[13],
                                   "step_share": d[12] / max(t_after[3] - t_before[3], 1e-9)} if r_g else None}

    # ---------------- e2e through the reference-facing entry points with pinned host arrays
    e2e = None
    if not args.no_e2e:
        nb = cfg.nbuffer
        pin = lambda *shape: torch.zeros(*shape, dtype=torch.float64).pin_memory().numpy()
        h_atype, h_q = pin(nb), pin(nb)
        h_pos, h_v, h_f = pin(3, nb), pin(3, nb), pin(3, nb)
        e.qs, e.qt, e.qsfp, e.qsfv = pin(nb), pin(nb), pin(nb), pin(nb)
        e.state_download(h_atype, h_pos, h_v, h_f, h_q)
        n = e.NATOMS
        mass = np.asarray(s.mass)
        ksteps = args.steps
        dthm_of_type = dt * 0.5 / np.maximum(mass, 1e-300)
        # the host integrator (the Fortran driver's O(N) loops, src/main.F90:64-72,86-98) as compiled loops
        import numba
        # one rank per GPU shares the host's cores: give each rank's integrator its share instead of a full-size thread pool
        ncores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
        numba.set_num_threads(max(1, min(numba.config.NUMBA_NUM_THREADS, ncores // max(world, 1))))

        @numba.njit(parallel=True, cache=False)
        def first_half(n, dt, lw2, dthm_t, atype, v, f, q, qsfp, qsfv, pos):
            for i in numba.prange(n):
                d = dthm_t[int(atype[i])]                  # int() == nint here: atype = type + gid*1e-13
                qsfv[i] += 0.5 * dt * lw2 * (q[i] - qsfp[i])
                qsfp[i] += dt * qsfv[i]
                for c in range(3):
                    v[c, i] += d * f[c, i]
                    pos[c, i] += dt * v[c, i]

        @numba.njit(parallel=True, cache=False)
        def second_half(n, dt, lw2, dthm_t, atype, v, f, q, qsfp, qsfv):
            for i in numba.prange(n):
                d = dthm_t[int(atype[i])]
                for c in range(3):
                    v[c, i] += d * f[c, i]
                qsfv[i] += 0.5 * dt * lw2 * (q[i] - qsfp[i])

        def host_step():
            nonlocal n
            first_half(n, dt, lw2, dthm_of_type, h_atype, h_v, h_f, h_q, e.qsfp, e.qsfv, h_pos)   # :64-72
            e.COPYATOMS(MODE_MOVE, [0.0, 0.0, 0.0], h_atype, h_pos, h_v, h_f, h_q)    # :75
            n = e.NATOMS
            e.QEq(h_atype, h_pos, h_q)                                                 # :80
            e.FORCE(h_atype, h_pos, h_f, h_q)                                          # :84
            second_half(n, dt, lw2, dthm_of_type, h_atype, h_v, h_f, h_q, e.qsfp, e.qsfv)         # :86-98
        host_step()
        barrier()
        tb = e.timers()
        t0 = time.perf_counter()
        for _ in range(ksteps):
            host_step()
        barrier()
        t_e2e = allmax(time.perf_counter() - t0)
        ta = e.timers()
        h2d, d2h = ta[20] - tb[20], ta[21] - tb[21]     # bytes the entry points actually copied (counted in the library)
        e2e = {"value": natoms_total * ksteps / t_e2e, "unit": "atom-timesteps/s", "h2d_bytes_per_step": int(h2d / ksteps),
               "d2h_bytes_per_step": int(d2h / ksteps), "steps": ksteps, "ms_per_step": t_e2e / ksteps * 1e3,
               "api": "Engine.COPYATOMS(MODE_MOVE) + Engine.QEq + Engine.FORCE over rxg_move/rxg_qeq/rxg_force, pinned host arrays, host integrator",
               "host_threads_per_rank": int(numba.get_num_threads())}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu, _ = cpu_sample(args, 20, 2)

    if rank == 0:
        line = {"metric": "atom-timesteps/s, RDX ReaxFF+QEq", "value": value, "unit": "atom-timesteps/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
                "scaling": "strong" if args.strong else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": f"RDX conf/init.rdx.lg (LG ffield) x{mc[0]}x{mc[1]}x{mc[2]} = {int(natoms_total)} atoms, ReaxFF+QEq "
                                       f"(QEq_tol 1e-7, NMAXQEq 500, every step), NVE dt {DT_FS} fs, gaussian displacements sigma={args.sigma} A",
                           "vprocs": list(vp), "atoms_per_gpu": nres, "parallelism": f"spatial decomposition {vp[0]}x{vp[1]}x{vp[2]}",
                           "ghost_refresh": ("peer-memory windows over NVLink (cudaIpc)" if (world > 1 and e.peer_halo()) else
                                             ("ncclSend/ncclRecv" if world > 1 else "periodic images, local gather")),
                           "cg_allreduce": ("peer-memory windows, rank-ordered sum" if (world > 1 and e.peer_allreduce()) else
                                            ("ncclAllReduce" if world > 1 else "none (1 rank)")),
                           "l2_policy": "inputs larger than L2 (QEq matrix ~5 GB per SpMV pass, >> 126 MB)",
                           "cg_iterations_per_step": cg_iters, "nnz": nnz, "pe_per_atom": pe[1:].sum() / max(e.natoms_resident(), 1)},
                "phase_ms_per_step": {"QEq": d[4] / args.steps, "FORCE": d[5] / args.steps, "MOVE": d[6] / args.steps},
                "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "clocks": clocks,
                "gpu_launches": int(l_after - l_before)}
        print(json.dumps(line), flush=True)
    e.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
