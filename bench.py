#!/usr/bin/env python
"""bench.py -- atom-timesteps/s of the ReaxFF+QEq hot path on N B200s (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W [--config rdx|water|sic|pqeq] [--strong]   # this framework, one rank per GPU
    python bench.py --impl reference --steps K --warmup W [--config ...]                    # CPU restatement of the reference

A "step" is one pass of the reference's main-loop body (src/main.F90:64-98): integrator halves, COPYATOMS(MOVE),
QEq / PQEq (CG to QEq_tol 1e-7) and FORCE over one synthetic replicated configuration with Gaussian sigma = 0.02 A
displacements (counter-based, seed 20261017), zero initial velocities and charges.

--config (rxmd_b200/host/configs.py; the headline is `rdx` = BASELINE.json configs[1], the one the metric is quoted on):
    rdx    conf/init.rdx.lg (LG ffield) x18^3 = 979 776 atoms per GPU
    water  conf/init.water ice Ih x(60,35,40) = 2 016 000 atoms          (configs[2]; north_star asks for --strong)
    sic    conf/init.sicnp x(20,20,18) = 3 938 400 atoms per GPU         (configs[3], weak scaling to vprocs 2 2 2)
    pqeq   conf/init.pe.pqeq + pqeq1.par x(30,45,88) = 1 425 600 atoms   (configs[4], PQEq, rctap 12.5 A)
N>1 is weak scaling (the same block per GPU, vprocs (2,1,1) (2,2,1) (2,2,2)); `--strong` keeps the block in total and
splits it over the ranks.  Every rank generates its own sub-domain only.

value : device-resident stepping (rxg_md_run), inputs in HBM when the clock starts, timed with CUDA events on the
        library's own stream, max over ranks.
e2e   : the same step driven through the reference-facing entry points COPYATOMS(MODE_MOVE) / QEq / FORCE with
        pinned HOST arrays (host<->device copies inside the timed region, integrator on the host like the Fortran
        driver).
parity: (N > 1) before timing, a ~10 k-atom run over the same multi-GPU data plane (NCCL exchange, peer-memory ghost refresh
        and all-reduce) is compared with the CPU oracle simulating the same vprocs: copyptr, 10 A row counts, forces 1e-9,
        energies 1e-9, charges, migration over 10 MD steps.
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
CPU_MC = {"rdx": (10, 10, 10), "water": (24, 14, 16), "sic": (6, 6, 6), "pqeq": (10, 15, 28)}        # cpu_baseline leg: 10-30 s
REF_MC = {"rdx": (12, 12, 13), "water": (32, 18, 20), "sic": (8, 8, 8), "pqeq": (14, 21, 40)}       # --impl reference: >= 300 k atoms


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


def host_cores():
    return len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)


def bind_rank_to_gpu_cores(local, world):
    """One rank per GPU on one node: give each rank the share of the host cores that sits next to ITS GPU (what a launch
    script does with numactl for the Fortran host).  The cores the driver reports as local to the GPU (NVML) are intersected
    with the cores this process may use and split evenly among the ranks whose GPUs share them; pinned buffers allocated
    afterwards land on that NUMA node.  Returns the cores taken, or None when the topology is not available."""
    if world <= 1 or not hasattr(os, "sched_setaffinity"):
        return None
    try:
        import pynvml
        pynvml.nvmlInit()
        allowed = sorted(os.sched_getaffinity(0))
        ncpu = os.cpu_count() or 1
        words = (ncpu + 63) // 64

        def local_set(dev):
            mask = pynvml.nvmlDeviceGetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(dev), words)
            cores = [64 * w + b for w, m in enumerate(mask) for b in range(64) if (int(m) >> b) & 1]
            return tuple(c for c in cores if c in set(allowed))
        sets = [local_set(d) for d in range(world)]
        mine = sets[local]
        if not mine:
            return None
        peers = [d for d in range(world) if sets[d] == mine]          # ranks whose GPUs hang off the same cores
        k = peers.index(local)
        share = [c for j, c in enumerate(mine) if j * len(peers) // len(mine) == k]
        if not share:
            return None
        os.sched_setaffinity(0, share)
        return share
    except Exception:
        return None


def cpu_sample(config, mc, steps, warmup, sigma, budget_s=None):
    """The CPU restatement of the reference (oracle/, OpenMP) on a bounded sample of the same workload, on all the host
    cores this process may use (the team size is set explicitly: launchers such as torch.distributed.run export
    OMP_NUM_THREADS=1).  With `budget_s` the number of timed steps shrinks so that the loop ends within the budget."""
    from rxmd_b200.host.configs import build_config
    from oracle.pyoracle import Oracle, set_threads
    threads = set_threads(host_cores())
    s, tot, vp, cfgkw, label = build_config(config, mc=mc, sigma=sigma)
    o = Oracle(s, s.config(**cfgkw))
    dt = DT_FS / UTIME
    lw2 = 2.0 * LEX_K / dt / dt
    o.move(); o.qeq(); o.force()
    t0 = time.perf_counter()
    o.md_run(warmup, dt, 1, lw2, 0)
    t_warm = (time.perf_counter() - t0) / max(warmup, 1)
    if budget_s is not None:
        steps = max(1, min(steps, int(budget_s / max(t_warm, 1e-9))))
    t0 = time.perf_counter()
    o.md_run(steps, dt, 1, lw2, warmup)
    t = time.perf_counter() - t0
    o.close()
    return {"value": s.natoms * steps / t, "unit": "atom-timesteps/s", "cores": threads, "kind": "port",
            "sample": f"{label} x{tot[0]}x{tot[1]}x{tot[2]} = {s.natoms} atoms, {steps} steps after {warmup} warm-up, sigma={sigma} A; "
                      f"OpenMP C++ restatement of the reference loops with {threads} threads (omp_get_max_threads; no Fortran "
                      f"toolchain in the image, oracle/_ref is empty)",
            "natoms": s.natoms, "steps": steps, "warmup": warmup}, t / steps * 1e3, label


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from rxmd_b200.host.configs import CONFIGS
    mc = tuple(args.cpu_mc) if args.cpu_mc else REF_MC[args.config]
    warm = max(1, min(args.warmup, 3))
    cb, ms, label = cpu_sample(args.config, mc, args.steps, warm, args.sigma, budget_s=150.0)
    full = CONFIGS[args.config]["mc"]
    line = {"impl": "reference", "metric": "atom-timesteps/s, RDX ReaxFF+QEq" if args.config == "rdx" else f"atom-timesteps/s, {label}",
            "value": cb["value"], "unit": "atom-timesteps/s",
            "n_gpus": args.gpus, "steps": cb["steps"], "warmup": warm, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"{label}, QEq_tol 1e-7 every step, NVE dt {DT_FS} fs", "sample": cb["sample"],
                       "same_config": list(mc) == list(full), "full_config_replication": list(full), "sample_replication": list(mc)},
            "cpu_baseline": cb, "e2e": {"value": cb["value"], "unit": "atom-timesteps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def preflight_parity(dist, rank, world, local):
    """~10 k atoms over the same multi-GPU data plane against the oracle with identical vprocs (tools/mr_diag.py)."""
    import torch
    from tools.mr_diag import compare
    mc = {2: (6, 3, 3), 4: (6, 6, 3), 8: (5, 5, 5)}[world]          # RDX 168-atom cells: 9 072 / 18 144 / 21 000 atoms
    res = compare(rank, world, local, mc, sigma=0.02, verbose=False)   # a failed comparison is reported in the line, never hidden
    flags = ["ok", "copyptr_qeq", "copyptr_force", "rows", "pe_ok", "migration", "nstep_same"]
    nums = ["dq", "f_rel", "md_pe_rel"]
    neg = lambda x: -x if x == x else -1e9          # NaN (a rank that raised) reads as a huge difference
    t = torch.tensor([1.0 if res.get(k) else 0.0 for k in flags] + [neg(float(res.get(k, float("nan")))) for k in nums],
                     dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    out = {"ok": bool(t[0].item() > 0.5), "ranks": world, "atoms": int(np.prod(mc)) * 168, "replication": list(mc)}
    for i, k in enumerate(flags[1:], start=1):
        out[k] = bool(t[i].item() > 0.5)
    out.update({"max_dq": -t[len(flags)].item(), "max_force_rel": -t[len(flags) + 1].item(), "md10_pe_rel": -t[len(flags) + 2].item(),
                "checked": "copyptr (QEq and FORCE halos), 10 A row counts, forces <= 1e-9, energies <= 1e-9 (pe_ok), charges (1e-6 same stop / 1e-4), "
                           "10 MD steps with migration: global PE <= 1e-6, atom counts", "peer_halo": res.get("peer_halo"),
                "peer_allreduce": res.get("peer_allreduce"), "error": res.get("error")})
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--config", default="rdx", choices=["rdx", "water", "sic", "pqeq"])
    ap.add_argument("--mc", type=int, nargs=3, default=None, help="unit-cell replication per GPU (default: the configuration's own)")
    ap.add_argument("--cpu-mc", type=int, nargs=3, default=None)
    ap.add_argument("--sigma", type=float, default=0.02)
    ap.add_argument("--strong", action="store_true", help="strong scaling: the replication is the TOTAL, split over the GPUs")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the multi-GPU parity pre-flight")
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
    bound = bind_rank_to_gpu_cores(local, world)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    from rxmd_b200.host.engine import Engine, MODE_MOVE, HINT_ATOMS_ON_DEVICE, HINT_Q_ON_DEVICE, HINT_DEFER_POS, HINT_CHARGES_STAY
    from rxmd_b200.host.configs import build_config
    # e2e leg: let rxg_force reuse the halo and 10 A list of the rxg_qeq that precedes it when the host hands back
    # bit-identical atoms (verified on the device); rxg_md_run does the same sharing internally
    parity = None
    if world > 1 and not args.no_parity:
        # literal entry points (QEq and FORCE each build their own halo and list, comparable array by array with the oracle);
        # the 10 MD steps of the comparison run rxg_md_run, which shares halo and list like the timed region does
        parity = preflight_parity(dist, rank, world, local)
    os.environ.setdefault("RXG_FUSE_API", "1")
    t_setup0 = time.perf_counter()
    s, mc, vp, cfgkw, label = build_config(args.config, mc=args.mc, nranks=world, strong=args.strong, sigma=args.sigma, only_rank=rank)
    t_setup = time.perf_counter() - t_setup0
    cfg = s.config(device=local, **cfgkw)
    pq = bool(cfg.isPQEq)
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

    def allred(x, op):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=op)
        return float(t.item())
    allmax = lambda x: allred(x, dist.ReduceOp.MAX)
    allsum = lambda x: allred(x, dist.ReduceOp.SUM)

    # ---------------- device-resident stepping: `value`
    atype, pos, v, f, q = e.host_arrays(st)
    e.state_upload(atype, pos, v, q)
    e.md_prime()
    e.md_run(args.warmup, dt, 1, lw2, 0)
    t_before = e.timers()
    it_before = e.it_timer()
    l_before = e.launches()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    e.md_run(args.steps, dt, 1, lw2, args.warmup)
    barrier()
    clocks = sampler.finish() if rank == 0 else None
    t_after = e.timers()
    it_after = e.it_timer()
    l_after = e.launches()
    ms_total = allmax(t_after[3] - t_before[3])
    # the reference's own phase table (it_timer slots, src/main.F90:135-180), CUDA-event times of rank 0, ms per step
    IT_NAMES = {3: "LINKEDLIST", 4: "COPYATOMS", 5: "NEIGHBORLIST", 6: "BOCALC", 7: "ENbond", 8: "Ebond", 9: "Elnpr", 10: "Ehb", 11: "E3b",
                12: "E4b", 13: "ForceBondedTerms", 15: "GetNonbondingPairList", 16: "qeq_initialize", 18: "get_hsh (CG)"}
    it_ms = {f"{k} {nm}": round((it_after[k - 1] - it_before[k - 1]) * 1e3 / max(args.steps, 1), 3) for k, nm in IT_NAMES.items()}
    natoms_total = allsum(float(e.natoms_resident()))
    value = natoms_total * args.steps / (ms_total * 1e-3)
    d = t_after - t_before
    cg_iters = d[17] / max(args.steps, 1)
    pe, ke, qsum, _ = e.md_observe()
    pe_global, ke_global, q_global = allsum(float(pe[1:].sum())), allsum(ke), allsum(qsum)    # global observables (PRINTE's all-reduce)

    # ---------------- roofline of the dominant kernel: the CG's sparse product
    nnz, nloc, ntot, nun = t_after[14], t_after[15], t_after[16], t_after[19]
    peak, peak_src = peak_hbm()
    traffic = None
    tp = os.path.join(ROOT, "profiles", "spmv_traffic.json")
    if os.path.exists(tp) and args.config == "rdx" and world == 1 and args.mc is None:
        try:
            traffic = json.load(open(tp)).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    roofline = None
    if d[11] > 0:
        bytes_alg = 12.0 * nnz + 4.0 * (nloc + 1) + 8.0 * 2 * ntot + 40.0 * nloc      # SURVEY 8(d), two gathered vectors (hs, ht)
        ach = bytes_alg / (d[10] / d[11] * 1e-3) / 1e9
        rows = os.environ.get("RXG_SPMV") != "items"
        win = (t_after[25] - t_before[25]) > 0 and (t_after[26] - t_before[26]) == 0
        roofline = {"kernel": ("k_spmv_win (QEq CG SpMV H.(hs,ht): x window of a cell group TMA-staged in shared memory, fp64 values + "
                               "16-bit window-relative columns streamed once, warp per row)" if win else
                               "k_spmv_rows (QEq CG SpMV H.(hs,ht): TMA-staged CSR stream, sub-warp per row)" if rows else
                               "k_spmv_items (QEq CG SpMV H.(hs,ht): cell-blocked union column stream + compacted fp64 values, "
                               "TMA ring, warp per row block)"),
                    "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": traffic,
                    "peak_source": peak_src, "algorithmic_bytes_per_launch": bytes_alg,
                    "stream_bytes_per_launch": ((10.0 * t_after[18] + 12.0 * nloc) if win else 12.0 * t_after[18] if rows else
                                                8.0 * t_after[18] + 5.0 * nun) + 16.0 * ntot + 32.0 * nloc,
                    "cells_per_group": int(t_after[27]) if win else None, "largest_window_entries": int(t_after[28]) if win else None,
                    "avg_launch_ms": d[10] / d[11], "launches_timed": int(d[11]),
                    "step_share": d[10] / max(t_after[3] - t_before[3], 1e-9)}
        # `frac` is quoted on the ALGORITHMIC bytes (12 B per entry); k_spmv_win moves 10 B per entry (16-bit columns),
        # so the fraction of the peak its actual stream reaches is stated next to it
        roofline["frac_on_bytes_moved"] = roofline["stream_bytes_per_launch"] / (roofline["avg_launch_ms"] * 1e-3) / 1e9 / peak

    # ---------------- e2e through the reference-facing entry points with pinned host arrays
    e2e = None
    if not args.no_e2e:
        nb = cfg.nbuffer
        pin = lambda *shape: torch.zeros(*shape, dtype=torch.float64).pin_memory().numpy()
        h_atype, h_q = pin(nb), pin(nb)
        h_pos, h_v, h_f = pin(3, nb), pin(3, nb), pin(3, nb)
        e.qs, e.qt, e.qsfp, e.qsfv = pin(nb), pin(nb), pin(nb), pin(nb)
        if pq:
            e.spos = pin(3, nb)
            e._chk(e.L.rxg_spos_download(e.h, e.natoms_resident(), e.spos.ctypes.data_as(e.L.rxg_spos_download.argtypes[2])))
        e.state_download(h_atype, h_pos, h_v, h_f, h_q)
        n = e.NATOMS
        mass = np.asarray(s.mass)
        ksteps = args.steps
        dthm_of_type = dt * 0.5 / np.maximum(mass, 1e-300)
        # the host integrator (the Fortran driver's O(N) loops, src/main.F90:64-72,86-98) as compiled loops
        import numba
        # one rank per GPU shares the host's cores: give each rank's integrator its share instead of a full-size thread pool
        numba.set_num_threads(max(1, min(numba.config.NUMBA_NUM_THREADS, len(bound) if bound else host_cores() // max(world, 1))))

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

        e2e_parts = np.zeros(5)    # seconds in: first half, COPYATOMS(MOVE), QEq, FORCE, second half (host clock, this rank)

        def host_step():
            nonlocal n
            t0 = time.perf_counter()
            # the hints are what rxmd_b200/gpu_shim.F90 states inside the main loop: nothing touches atype/pos/q between
            # COPYATOMS(MOVE), QEq and FORCE (src/main.F90:75-84), so pos travels up once and down once per step
            first_half(n, dt, lw2, dthm_of_type, h_atype, h_v, h_f, h_q, e.qsfp, e.qsfv, h_pos)   # :64-72
            t1 = time.perf_counter()
            e.hint(HINT_DEFER_POS | HINT_CHARGES_STAY)
            e.COPYATOMS(MODE_MOVE, [0.0, 0.0, 0.0], h_atype, h_pos, h_v, h_f, h_q)    # :75
            n = e.NATOMS
            t2 = time.perf_counter()
            e.hint(HINT_ATOMS_ON_DEVICE | HINT_Q_ON_DEVICE | HINT_DEFER_POS)
            if pq:
                e.PQEq(h_atype, h_pos, h_q)                                            # :78
            else:
                e.QEq(h_atype, h_pos, h_q)                                             # :80
            t3 = time.perf_counter()
            e.hint(HINT_ATOMS_ON_DEVICE | HINT_Q_ON_DEVICE)
            e.FORCE(h_atype, h_pos, h_f, h_q)                                          # :84
            t4 = time.perf_counter()
            second_half(n, dt, lw2, dthm_of_type, h_atype, h_v, h_f, h_q, e.qsfp, e.qsfv)         # :86-98
            e2e_parts[:] += (t1 - t0, t2 - t1, t3 - t2, t4 - t3, time.perf_counter() - t4)
        host_step()
        barrier()
        e2e_parts[:] = 0.0
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
               "api": f"Engine.COPYATOMS(MODE_MOVE) + Engine.{'PQEq' if pq else 'QEq'} + Engine.FORCE over rxg_move/rxg_{'pqeq' if pq else 'qeq'}/rxg_force, "
                      "pinned host arrays, host integrator",
               "host_threads_per_rank": int(numba.get_num_threads()), "host_cores_bound_to_gpu": (len(bound) if bound else None),
               "force_calls_reusing_qeq_list": int(ta[22] - tb[22]),
               "ms_per_step_by_call": dict(zip(["host first half", "COPYATOMS(MOVE)", "QEq", "FORCE", "host second half"],
                                               [round(float(x) / ksteps * 1e3, 3) for x in e2e_parts]))}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu, _, _ = cpu_sample(args.config, tuple(args.cpu_mc) if args.cpu_mc else CPU_MC[args.config], 8, 1, args.sigma, budget_s=25.0)

    if rank == 0:
        line = {"metric": "atom-timesteps/s, RDX ReaxFF+QEq" if args.config == "rdx" else f"atom-timesteps/s, {label}",
                "value": value, "unit": "atom-timesteps/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
                "scaling": "strong" if args.strong else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": f"{label} x{mc[0]}x{mc[1]}x{mc[2]} = {int(natoms_total)} atoms, "
                                       f"(QEq_tol 1e-7, NMAXQEq 500, every step), NVE dt {DT_FS} fs, gaussian displacements sigma={args.sigma} A",
                           "name": args.config, "vprocs": list(vp), "atoms_per_gpu": nres, "parallelism": f"spatial decomposition {vp[0]}x{vp[1]}x{vp[2]}",
                           "ghost_refresh": ("peer-memory windows over NVLink (cudaIpc)" if (world > 1 and e.peer_halo()) else
                                             ("ncclSend/ncclRecv" if world > 1 else "periodic images, local gather")),
                           "cg_allreduce": ("peer-memory windows, rank-ordered sum" if (world > 1 and e.peer_allreduce()) else
                                            ("ncclAllReduce" if world > 1 else "none (1 rank)")),
                           "l2_policy": "inputs larger than L2 (QEq matrix ~4 GB per SpMV pass, >> 126 MB)",
                           "cg_iterations_per_step": cg_iters, "nnz": nnz, "union_entries": nun,
                           "list_builds_without_count_pass": int(d[23]), "of_those_rebuilt_after_row_overflow": int(d[24]),
                           "pe_per_atom_global": pe_global / max(natoms_total, 1), "ke_per_atom_global": ke_global / max(natoms_total, 1),
                           "sum_q_global": q_global, "setup_seconds_per_rank": t_setup},
                "phase_ms_per_step": {"QEq": d[4] / args.steps, "FORCE": d[5] / args.steps, "MOVE": d[6] / args.steps},
                "it_timer_ms_per_step": it_ms,
                "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "clocks": clocks, "parity": parity,
                "gpu_launches": int(l_after - l_before)}
        print(json.dumps(line), flush=True)
    e.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
