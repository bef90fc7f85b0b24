"""Run md_prime + a few device-resident steps of the bench workload (for ncu captures)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import argparse
import bench
from rxmd_b200.host.engine import Engine

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="rdx")
ap.add_argument("--mc", type=int, nargs=3, default=None)
ap.add_argument("--steps", type=int, default=2)
ap.add_argument("--sigma", type=float, default=0.02)
a = ap.parse_args()
from rxmd_b200.host.configs import build_config
s, mc, vp, cfgkw, label = build_config(a.config, mc=a.mc, sigma=a.sigma)
cfg = s.config(**cfgkw)
e = Engine(s, cfg)
atype, pos, v, f, q = e.host_arrays(s.ranks[0])
e.state_upload(atype, pos, v, q)
dt = bench.DT_FS / bench.UTIME
e.md_prime()
e.md_run(a.steps, dt, 1, 2.0 * bench.LEX_K / dt / dt, 0)
t = e.timers()
print("ms/step", t[3] / a.steps, "QEq", t[4] / a.steps, "FORCE", t[5] / a.steps, "cg iters", t[17], "launches", e.launches())
e.close()
