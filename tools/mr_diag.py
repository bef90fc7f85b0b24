"""Multi-rank GPU-vs-oracle comparison (development diagnostic).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        tools/mr_diag.py mcx mcy mcz [--sigma S]

Every rank builds the same system, drives its own GPU through the C-ABI, and compares its arrays with the same rank
of the oracle (which simulates all ranks of the identical `vprocs` decomposition in-process).
"""
import os
import sys
import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rxmd_b200.host.system import build_system  # noqa: E402
from rxmd_b200.host.engine import Engine  # noqa: E402
from oracle.pyoracle import Oracle  # noqa: E402

G = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests/golden/inputs/init.rdx/")
VP = {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}


def rel(a, b):
    d = np.abs(a - b).max() if a.size else 0.0
    return d, d / max(np.abs(b).max() if b.size else 0.0, 1e-300)


def _compare_body(rank, world, local, mc, sigma, pqeq, holder):
    """One multi-rank comparison over the library's own data plane (NCCL exchange in MODE_COPY/MOVE/CPBK, peer-memory windows
    for the per-iteration refreshes and all-reduces) against the oracle simulating the same `vprocs`.  Needs an initialised
    torch.distributed NCCL group.  Returns a dict of this rank's findings (all-reduced into a verdict by the caller)."""
    vp = VP[world]
    from oracle.pyoracle import set_threads
    ncores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    set_threads(max(1, ncores // world))      # every rank runs its own copy of the oracle: share the host's cores
    if pqeq:
        GP = G.replace("init.rdx/", "init.pe.pqeq/")
        s = build_system(GP + "input.xyz", GP + "ffield", mc=mc, vprocs=vp, displace_sigma=sigma, pqeq_path=GP + "pqeq1.par")
    else:
        s = build_system(G + "input.xyz", G + "ffield", mc=mc, vprocs=vp, displace_sigma=sigma)
    cfg = s.config(device=local)
    e = Engine(s, cfg, rank=rank)
    holder["e"] = e
    e.comm_init_torch(dist)
    o = Oracle(s, cfg)       # all ranks, simulated
    holder["o"] = o
    atype, pos, v, f, q = e.host_arrays(s.ranks[rank])
    n = e.NATOMS
    out = [f"rank {rank}/{world} vprocs {vp} natoms {n} of {s.natoms} peer_halo {e.peer_halo()}"]
    if pqeq:
        for r in range(world):
            nr = len(s.ranks[r]["atype"])
            sp = np.random.default_rng(100 + r).normal(0.0, 4e-3, (3, nr))
            o.set_spos(r, sp)
            if r == rank:
                e.spos[:, :n] = sp
    o.qeq()
    if pqeq:
        e.PQEq(atype, pos, q)
    else:
        e.QEq(atype, pos, q)
    cp_o, cp_g = o.i32("copyptr", rank), e.fetch("copyptr")
    out.append(f"  QEq copyptr equal {np.array_equal(cp_o, cp_g)} {cp_g.tolist()} nstep {o.observe()[3]} vs {e.nstep_qeq}")
    if np.array_equal(cp_o, cp_g):
        out.append(f"  ghost pos diff {rel(e.fetch('pos').reshape(3, -1), o.f64('pos', rank).reshape(3, -1))} atype equal "
                   f"{np.array_equal(e.fetch('atype'), o.f64('atype', rank))}")
    rb, re_ = e.fetch("rowbeg"), e.fetch("rowend")
    rows_ok = np.array_equal(o.i32('nbpcnt', rank), re_ - rb)
    cp_qeq_ok = np.array_equal(cp_o, cp_g)
    out.append(f"  row counts equal {rows_ok}")
    out.append(f"  q diff {rel(q[:n], o.f64('q', rank)[:n])}")
    dq = rel(q[:n], o.f64('q', rank)[:n])[0]
    nstep_same = o.observe()[3] == e.nstep_qeq
    q[:n] = o.f64("q", rank)[:n]
    if pqeq:
        sp_o = o.f64("spos", rank).reshape(3, -1)[:, :n]
        out.append(f"  spos diff {rel(e.spos[:, :n], sp_o)} skips {e.pqeq_skips()} vs {o.i32('pqeq_skips', rank)[0]}")
        e.spos[:, :n] = sp_o
        e._chk(e.L.rxg_spos_upload(e.h, n, e.spos.ctypes.data_as(e.L.rxg_spos_upload.argtypes[2])))
    o.force()
    e.FORCE(atype, pos, f, q)
    cp_o, cp_g = o.i32("copyptr", rank), e.fetch("copyptr")
    out.append(f"  FORCE copyptr equal {np.array_equal(cp_o, cp_g)} {cp_g.tolist()}")
    pe_o = o.f64("PE", rank)
    out.append(f"  PE rel diff max {np.max(np.abs(e.PE[1:] - pe_o[1:]) / np.maximum(np.abs(pe_o[1:]), 1e-300))}")
    pe_ok = np.max(np.abs(e.PE[1:] - pe_o[1:]) / np.maximum(np.abs(pe_o[1:]), 1e-6 * np.abs(pe_o[1:]).max())) < 1e-9
    cp_force_ok = np.array_equal(cp_o, cp_g)
    f_o = o.f64("f", rank).reshape(3, -1)[:, :n]
    out.append(f"  f diff {rel(f[:, :n], f_o)}")
    # device-resident MD with migration
    UTIME = 1e3 / 20.455
    dt = 0.25 / UTIME
    lw2 = 2.0 * 2.0 / dt / dt
    os.environ.setdefault("X", "")
    if pqeq:   # fast atoms so that some cross the rank boundary and carry their shells along
        vv = np.random.default_rng(5).normal(0.0, 2e-2, (3, s.natoms))
        off = 0
        for r in range(world):
            nr = len(s.ranks[r]["atype"])
            st = s.ranks[r]
            o.set_atoms(r, st["atype"], st["pos"], vv[:, off:off + nr].copy(), o.f64("q", r)[:nr].copy())
            o.set_spos(r, o.f64("spos", r).reshape(3, -1)[:, :nr].copy())
            if r == rank:
                v[:, :n] = vv[:, off:off + nr]
                pos[:, :n] = st["pos"]
            off += nr
    e.state_upload(atype, pos, v, q)
    e.md_prime()
    o.qeq(); o.force()
    e.md_run(10, dt, 1, lw2, 0)
    o.md_run(10, dt, 1, lw2, 0)
    pe_g, ke_g, qs_g, it_g = e.md_observe()
    tot = torch.tensor([pe_g[1:].sum(), ke_g, float(e.natoms_resident())], dtype=torch.float64, device="cuda")
    dist.all_reduce(tot)
    pe_oa, ke_o, _, it_o = o.observe()
    out.append(f"  MD10: natoms {e.natoms_resident()} vs {o.natoms(rank)}  global PE {tot[0].item():.9f} vs {pe_oa[0]:.9f}  KE {tot[1].item():.9e} vs {ke_o:.9e} "
               f"natoms_total {int(tot[2].item())}")
    res = {"copyptr_qeq": bool(cp_qeq_ok), "copyptr_force": bool(cp_force_ok), "rows": bool(rows_ok), "nstep_same": bool(nstep_same),
           "dq": float(dq), "f_rel": float(rel(f[:, :n], f_o)[1]), "pe_ok": bool(pe_ok), "peer_halo": bool(e.peer_halo()),
           "peer_allreduce": bool(e.peer_allreduce()),
           "migration": bool(e.natoms_resident() == o.natoms(rank) and int(tot[2].item()) == s.natoms),
           "md_pe_rel": float(abs(tot[0].item() - pe_oa[0]) / abs(pe_oa[0]))}
    # charges: the production CG's bars (tests/test_gpu_parity.py: 1e-6 when both sides stop in the same iteration, 1e-4 otherwise)
    res["ok"] = bool(res["copyptr_qeq"] and res["copyptr_force"] and res["rows"] and res["f_rel"] < 1e-9 and res["pe_ok"] and
                     res["migration"] and res["md_pe_rel"] < 1e-6 and res["dq"] <= (1e-6 if nstep_same else 1e-4))
    return res, out


def compare(rank, world, local, mc, sigma=0.0, pqeq=False, do_assert=False, verbose=True):
    """Collective: every rank reaches the closing barrier and tears its engine down even if its own comparison raised, so
    that a failure is reported instead of leaving the other ranks waiting."""
    holder, out = {}, []
    try:
        res, out = _compare_body(rank, world, local, mc, sigma, pqeq, holder)
    except Exception as ex:
        res = {"ok": False, "error": repr(ex)[:300], "dq": float("nan"), "f_rel": float("nan"), "md_pe_rel": float("nan"),
               "nstep_same": False, "peer_halo": False, "peer_allreduce": False}
    if verbose:
        for r in range(world):
            dist.barrier()
            if r == rank:
                print("\n".join(out), flush=True)
    try:
        torch.cuda.synchronize()
    except Exception:
        pass
    dist.barrier()           # nobody frees its peer window while a neighbour may still be writing into it
    if "e" in holder:
        holder["e"].close()
    if "o" in holder:
        holder["o"].close()
    if do_assert:
        assert res["ok"], res
    return res


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    mc = tuple(int(x) for x in args[:3])
    sigma = float(sys.argv[sys.argv.index("--sigma") + 1]) if "--sigma" in sys.argv else 0.0
    res = compare(rank, world, local, mc, sigma, pqeq="--pqeq" in sys.argv, do_assert="--assert" in sys.argv)
    if rank == 0:
        print("rank 0 verdict:", res, flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
