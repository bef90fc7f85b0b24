"""Production CG vs the oracle: charge differences per system, with the stop rule and with a fixed number of iterations.

    python tools/cg_spread.py            # run on a B200 (gpurun)

Prints, per parity system: nstep_qeq on both sides, max |dq| at QEq_tol 1e-7, and max |dq| after exactly k iterations
(NMAXQEq = k, k = 1..6) -- the numbers behind the charge bars of tests/test_gpu_parity.py."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from rxmd_b200.host.system import build_system
from rxmd_b200.host.engine import Engine
from oracle.pyoracle import Oracle
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from test_gpu_parity import systems   # noqa: E402


def one(name, **cfgkw):
    kw = dict(systems()[name])
    s = build_system(kw.pop("xyz"), kw.pop("ff"), **kw)
    cfg = s.config(**cfgkw)
    e, o = Engine(s, cfg), Oracle(s, cfg)
    atype, pos, v, f, q = e.host_arrays(s.ranks[0])
    n = e.NATOMS
    o.qeq(); e.QEq(atype, pos, q)
    d = np.abs(q[:n] - o.f64("q")[:n]).max()
    r = (e.nstep_qeq, o.observe()[3], d, np.abs(o.f64("q")[:n]).max())
    e.close(); o.close()
    return r


for name in systems():
    g, oo, d, qm = one(name)
    print(f"{name:20s} tol 1e-7: nstep gpu {g} oracle {oo} max|dq| {d:.3e} (max|q| {qm:.3f})", flush=True)
    for k in (1, 2, 3, 4, 6, 10, 20):
        g, oo, d, qm = one(name, NMAXQEq=k)
        print(f"{'':20s} NMAXQEq={k}: nstep {g}/{oo} max|dq| {d:.3e}", flush=True)

# PQEq (polyethylene, pqeq1.par, rctap 12.5 A): the same comparison
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import test_gpu_pqeq as tp   # noqa: E402
for shell_sigma in (0.0, 4e-3):
    for kw in ({}, {"NMAXQEq": 1}, {"NMAXQEq": 2}, {"NMAXQEq": 4}, {"NMAXQEq": 6}, {"NMAXQEq": 10}, {"NMAXQEq": 20}):
        s, cfg, e, o, sp = tp.make(shell_sigma=shell_sigma, **kw)
        atype, pos, v, f, q = e.host_arrays(s.ranks[0])
        n = e.NATOMS
        e.spos[:, :n] = sp
        o.set_spos(0, sp)
        o.qeq(); e.PQEq(atype, pos, q)
        d = np.abs(q[:n] - o.f64("q")[:n]).max()
        ds = np.abs(e.spos[:, :n] - o.f64("spos").reshape(3, -1)[:, :n]).max()
        print(f"pqeq shell_sigma={shell_sigma} {kw}: nstep gpu {e.nstep_qeq} oracle {o.i32('nstep_qeq')[0]} max|dq| {d:.3e} max|dspos| {ds:.3e}", flush=True)
        e.close(); o.close()
