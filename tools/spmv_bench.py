"""Time the CG's sparse product on the bench workload for several kernel variants (one process, one system build).

    python tools/spmv_bench.py [--mc 18 18 18] [--config rdx|water|sic|pqeq] [--reps 20] variant[,variant...]

A variant is a comma-free list of KEY=VALUE pairs joined by '+', e.g.  RXG_SPMV=rows  or  RXG_SPMV_LEAD=2 ; 'default' = none.
Prints per variant the average launch time (CUDA events over `reps` back-to-back launches), the canonical-bytes GB/s
(12 nnz + 4 (N+1) + 16 (N+G) + 40 N) and the bytes the kernel's streams actually hold.
"""
import argparse
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from rxmd_b200.host.engine import Engine
from rxmd_b200.host.configs import build_config


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="rdx")
    ap.add_argument("--mc", type=int, nargs=3, default=None)
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("variants", nargs="*", default=["default"])
    a = ap.parse_args()
    s, tot, vp, cfgkw, label = build_config(a.config, mc=a.mc)
    keys = set()
    for v in a.variants:
        env = {} if v == "default" else dict(kv.split("=") for kv in v.split("+"))
        for k in keys:
            os.environ.pop(k, None)
        os.environ.update(env)
        keys |= set(env)
        cfg = s.config(NMAXQEq=2, **cfgkw)
        e = Engine(s, cfg)
        atype, pos, vv, f, q = e.host_arrays(s.ranks[0])
        if cfg.isPQEq:
            e.PQEq(atype, pos, q)
        else:
            e.QEq(atype, pos, q)
        t = e.timers()
        nnz, n, ntot, nun = t[14], t[15], t[16], t[19]
        x = np.random.default_rng(1).normal(0.0, 1.0, (int(ntot), 2))
        _, ms = e.debug_spmv(x, reps=a.reps)
        canon = 12.0 * nnz + 4.0 * (n + 1) + 16.0 * ntot + 40.0 * n
        t2 = e.timers()
        win = t2[25] > 0
        held = 8.0 * t[18] + (5.0 * nun if os.environ.get("RXG_SPMV") == "items" else (2.0 if win else 4.0) * t[18])
        if win:
            v += f" [win G={int(t2[27])} wmax={int(t2[28])}]"
        print(f"{a.config} {v:40s} {ms:8.4f} ms  canonical {canon / 1e9:.3f} GB -> {canon / ms / 1e6:8.1f} GB/s   streams {held / 1e9:.3f} GB -> "
              f"{held / ms / 1e6:8.1f} GB/s   (nnz {int(nnz)}, union {int(nun)}, ratio {nun / max(nnz, 1):.3f})", flush=True)
        e.close()


if __name__ == "__main__":
    main()
