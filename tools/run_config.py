"""Run one of the BASELINE.json configurations on one GPU: QEq + FORCE + a few device-resident steps (sanity/timing)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from rxmd_b200.host.system import build_system
from rxmd_b200.host.engine import Engine

INP = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "inputs")
CONFIGS = {
    "rdx1m": dict(d="init.rdx.lg", xyz="input.xyz", mc=(18, 18, 18), isLG=True),
    "water2m": dict(d="init.water", xyz="ice-1h.xyz", mc=(60, 35, 40), real_coords=True),
    "sic4m": dict(d="init.sicnp", xyz="input.xyz", mc=(20, 20, 18)),
    "pe1m": dict(d="init.pe.pqeq", xyz="input.xyz", mc=(30, 45, 88), pqeq="pqeq1.par"),     # BASELINE config 5 (PQEq, rctap 12.5 A)
    "pe_small": dict(d="init.pe.pqeq", xyz="input.xyz", mc=(10, 15, 30), pqeq="pqeq1.par"),
}
name = sys.argv[1]
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
kw = dict(CONFIGS[name])
d = kw.pop("d"); xyz = kw.pop("xyz")
if "pqeq" in kw:
    kw["pqeq_path"] = os.path.join(INP, d, kw.pop("pqeq"))
t0 = time.time()
s = build_system(os.path.join(INP, d, xyz), os.path.join(INP, d, "ffield"), displace_sigma=0.02, **kw)
cfg = s.config(maxneighbs10=2400 if s.pqeq is not None else 1500)
print(name, "natoms", s.natoms, "nbuffer", cfg.nbuffer, "maxrc", round(s.maxrc, 3), "build", round(time.time() - t0, 1), "s", flush=True)
e = Engine(s, cfg)
atype, pos, v, f, q = e.host_arrays(s.ranks[0])
e.state_upload(atype, pos, v, q)
t0 = time.time(); e.md_prime(); print("prime (QEq+FORCE) s", round(time.time() - t0, 3))
pe, ke, qs, it = e.md_observe()
print("PE/atom", pe[1:].sum() / s.natoms, "terms/atom", np.round(pe[1:] / s.natoms, 5), "nstep_qeq", it, "sum q", qs)
UTIME = 1e3 / 20.455
dt = 0.25 / UTIME
t0 = e.timers()
e.md_run(steps, dt, 1, 4.0 / dt / dt, 0)
t1 = e.timers()
dt_ = t1 - t0
pe, ke, qs, it = e.md_observe()
print(f"{steps} steps: {dt_[3] / steps:.2f} ms/step (QEq {dt_[4] / steps:.2f} FORCE {dt_[5] / steps:.2f}) => {s.natoms * steps / (dt_[3] * 1e-3) / 1e6:.2f} M atom-steps/s;"
      f" CG it/step {dt_[17] / steps:.1f}; SpMV {dt_[10] / max(dt_[11], 1):.3f} ms; nnz {t1[14]:.3e}; PE/atom {pe[1:].sum() / s.natoms:.6f} KE/atom {ke / s.natoms:.3e}")
e.close()
