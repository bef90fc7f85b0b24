"""md_prime + a few device-resident PQEq steps of BASELINE config 5 (polyethylene, rctap 12.5 A) for ncu launch lists."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rxmd_b200.host.system import build_system
from rxmd_b200.host.engine import Engine
INP = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "inputs", "init.pe.pqeq")
mc = tuple(int(x) for x in sys.argv[1:4]) if len(sys.argv) > 3 else (30, 45, 88)
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 2
s = build_system(os.path.join(INP, "input.xyz"), os.path.join(INP, "ffield"), mc=mc, displace_sigma=0.02, pqeq_path=os.path.join(INP, "pqeq1.par"))
e = Engine(s, s.config(maxneighbs10=2400))
atype, pos, v, f, q = e.host_arrays(s.ranks[0])
e.state_upload(atype, pos, v, q)
e.md_prime()
t0 = e.timers()
e.md_run(steps, 0.25 / (1e3 / 20.455), 1, 0.0, 0)
t = e.timers() - t0
print(f"natoms {s.natoms} ms/step {t[3] / steps:.2f} QEq {t[4] / steps:.2f} FORCE {t[5] / steps:.2f} cg/step {t[17] / steps:.1f} spmv {t[10] / max(t[11], 1):.3f} ms nnz {e.timers()[14]:.3e} skips {e.pqeq_skips()}")
e.close()
