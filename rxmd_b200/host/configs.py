"""The BASELINE.json configurations as synthetic replicated geometries: one definition for bench.py, the
profiling tools and the tests.  `mc` is the unit-cell replication PER GPU for weak scaling; `build_config(..., strong=True)`
keeps it as the total."""
from __future__ import annotations

import os

from .system import build_system

INPUTS = os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tests", "golden", "inputs")
VPROCS = {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}

CONFIGS = {
    # BASELINE configs[1]: conf/init.rdx.lg (LG force field) x18^3 = 979 776 atoms
    "rdx": dict(d="init.rdx.lg", xyz="input.xyz", mc=(18, 18, 18), isLG=True,
                label="RDX conf/init.rdx.lg (LG ffield) ReaxFF+QEq"),
    # configs[2]: conf/init.water ice Ih x(60,35,40) = 2 016 000 atoms (Ehb == 0 with the shipped ffield)
    "water": dict(d="init.water", xyz="ice-1h.xyz", mc=(60, 35, 40), real_coords=True,
                  label="water conf/init.water (ice Ih) ReaxFF+QEq"),
    # configs[3]: conf/init.sicnp x(20,20,18) = 3 938 400 atoms per GPU
    "sic": dict(d="init.sicnp", xyz="input.xyz", mc=(20, 20, 18), label="SiC nanoparticles conf/init.sicnp ReaxFF+QEq"),
    # configs[4]: conf/init.pe.pqeq with examples/3-reaxpq+/pqeq1.par x(30,45,88) = 1 425 600 atoms, rctap 12.5 A
    "pqeq": dict(d="init.pe.pqeq", xyz="input.xyz", mc=(30, 45, 88), pqeq="pqeq1.par", maxneighbs10=2400,
                 label="polyethylene conf/init.pe.pqeq ReaxFF+PQEq (pqeq1.par, rctap 12.5 A)"),
}


def build_config(name, mc=None, nranks=1, strong=False, sigma=0.02, only_rank=None):
    """-> (System, total replication, vprocs, cfg_kwargs, label)."""
    c = dict(CONFIGS[name])
    vp = VPROCS[nranks]
    if os.environ.get("RXG_BENCH_VPROCS"):      # diagnostics: another decomposition of the same rank count, e.g. "1,1,2"
        vp = tuple(int(x) for x in os.environ["RXG_BENCH_VPROCS"].split(","))
        assert vp[0] * vp[1] * vp[2] == nranks
    mc = tuple(mc) if mc is not None else c["mc"]
    if strong:
        tot = mc          # atoms go to ranks by position (init/geninit.F90:495-500): the replication need not divide by vprocs
    else:
        tot = tuple(mc[a] * vp[a] for a in range(3))
    d = os.path.join(INPUTS, c["d"])
    kw = {}
    if c.get("isLG"):
        kw["isLG"] = True
    if c.get("real_coords"):
        kw["real_coords"] = True
    if c.get("pqeq"):
        kw["pqeq_path"] = os.path.join(d, c["pqeq"])
    s = build_system(os.path.join(d, c["xyz"]), os.path.join(d, "ffield"), mc=tot, vprocs=vp, displace_sigma=sigma,
                     only_rank=only_rank, **kw)
    cfgkw = {"maxneighbs10": c["maxneighbs10"]} if "maxneighbs10" in c else {}
    return s, tot, vp, cfgkw, c["label"]
