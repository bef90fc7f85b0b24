"""Trace of the QEq CG iteration count per MD step (the workload definition behind atom-timesteps/s)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from rxmd_b200.host.system import build_system
from rxmd_b200.host.engine import Engine
INP = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "inputs", "init.rdx.lg")
mc = tuple(int(x) for x in sys.argv[1:4]) if len(sys.argv) > 3 else (18, 18, 18)
nsteps = int(sys.argv[4]) if len(sys.argv) > 4 else 60
s = build_system(os.path.join(INP, "input.xyz"), os.path.join(INP, "ffield"), mc=mc, isLG=True, displace_sigma=0.02)
e = Engine(s, s.config())
atype, pos, v, f, q = e.host_arrays(s.ranks[0])
e.state_upload(atype, pos, v, q)
e.md_prime()
UTIME = 1e3 / 20.455
dt = 0.25 / UTIME
its, ms = [e.md_observe()[3]], []
for k in range(nsteps):
    t0 = e.timers()
    e.md_run(1, dt, 1, 0.0, k)
    t1 = e.timers()
    pe, ke, qs, it = e.md_observe()
    its.append(it); ms.append(round(t1[3] - t0[3], 1))
print("natoms", s.natoms, "CG iterations per step:", its)
print("ms per step:", ms)
print("TE/atom drift:", (pe[1:].sum() + ke) / s.natoms)
e.close()
