"""Assemble one simulation's host-side inputs (the job of GETPARAMS + INITSYSTEM + geninit).

`build_system` returns everything the hot-path library (or the test oracle) needs for every rank
of a `vprocs` decomposition: packed force field, per-rank box structs and per-rank resident atoms
in REAL coordinates (ReadBIN's xs2xu, reference src/fileio.F90:536).
"""
from __future__ import annotations

from dataclasses import dataclass, field
import os
import numpy as np

from . import setup as S
from .ffield import read_ffield
from .geninit import read_xyz, replicate, hash_normal
from .binding import PackedFF, PackedBox, RxgConfig


@dataclass
class System:
    ff: object
    pff: PackedFF
    boxes: list            # PackedBox per rank
    vprocs: tuple
    lattice: tuple
    ranks: list            # per rank: dict(atype[n], pos[3,n] real, v[3,n], q[n])
    natoms: int
    maxrc: float
    rctap: float
    mass: np.ndarray       # per type, 1-based
    cfg_defaults: dict = field(default_factory=dict)
    pqeq: object = None    # setup.PQEqParams when built with a PQEq parameter file (--pqeq / PQEqParm, src/cmdline.F90:111-127)

    def config(self, nbuffer=None, device=0, isQEq=1, NMAXQEq=500, QEq_tol=1e-7, maxneighbs=30,
               maxneighbs10=1500, nmincell=S.NMINCELL, Lex_fqs=1.0, efield=None):
        c = RxgConfig()
        if nbuffer is None:
            nmax = max(len(r["atype"]) for r in self.ranks if r is not None)
            nbuffer = estimate_nbuffer(self, nmax)
        c.device, c.nbuffer, c.maxneighbs, c.maxneighbs10, c.nmincell = device, int(nbuffer), maxneighbs, maxneighbs10, nmincell
        c.isQEq, c.NMAXQEq, c.isPQEq, c.isEfield, c.eFieldDir = isQEq, NMAXQEq, int(self.pqeq is not None), 0, 1
        c.QEq_tol, c.Lex_fqs, c.eFieldStrength = QEq_tol, Lex_fqs, 0.0
        if efield is not None:                       # rxmd.in `efield <dir> <strength>`, src/cmdline.F90:287-290
            c.isEfield, c.eFieldDir, c.eFieldStrength = 1, int(efield[0]), float(efield[1])
        return c


def estimate_nbuffer(sysm, nres):
    """Residents + ghosts of the widest halo (FORCE: NMINCELL cells; QEq: rctap), with head-room."""
    b = sysm.boxes[0].struct
    lat = (b.lata, b.latb, b.latc)
    fr = 1.0
    for a in range(3):
        halo_f = S.NMINCELL * b.lcsize[a] / b.LBOX[a]
        halo_q = sysm.rctap / lat[a] / b.LBOX[a]
        fr *= 1.0 + 2.0 * max(halo_f, halo_q)
    return int(nres * fr * 1.08) + 1024


def build_system(xyz_path, ffield_path, mc=(1, 1, 1), vprocs=(1, 1, 1), isLG=False, real_coords=False,
                 displace_sigma=0.0, seed=20261017, pqeq_path=None, only_rank=None):
    """`only_rank=r`: generate rank r's atoms alone (`ranks[r']` is None for the others) -- what one process of a multi-GPU
    run needs; time and memory then scale with that rank's share."""
    ff = read_ffield(ffield_path, isLG=isLG)
    types0, pos0, lat0 = read_xyz(xyz_path, ff.atmname, real_coords=real_coords)
    displace = None
    if displace_sigma > 0.0:
        Hbig = S.get_box_params(lat0[0] * mc[0], lat0[1] * mc[1], lat0[2] * mc[2], lat0[3], lat0[4], lat0[5])
        Hbig_i = np.linalg.inv(Hbig)
        # Gaussian displacements indexed by global atom id (counter-based), so every decomposition sees the same geometry
        displace = lambda gid: (displace_sigma * hash_normal(seed, gid)) @ Hbig_i.T
    gen = replicate(types0, pos0, lat0, mc, vprocs, displace=displace, only_rank=only_rank)
    lattice = gen["lattice"]
    rctap = S.RCTAP0_PQEQ if pqeq_path else S.RCTAP0   # src/init.F90:28-32
    CTap = S.taper(rctap)
    pq, chi, eta = None, None, None
    if pqeq_path:                                     # get_pqeq_parms + initialize_pqeq, src/init.F90:40-41
        pq = S.read_pqeq_parms(pqeq_path)
        if pq.ntype > ff.nso:
            # initialize_pqeq writes chi(1:ntype_pqeq), eta(1:ntype_pqeq) of arrays sized nso (src/module.F90:492,517-523): a
            # parameter file with more rows than the force field has elements (conf/init.pe.pqeq/pqeq.in: 9 vs 7) overruns
            # them in the reference.  Rows past nso can never be addressed by an atom type, so they are dropped here.
            S.truncate_pqeq(pq, ff.nso)
        chi_full, eta_full = np.array(ff.chi, dtype=float), np.array(ff.eta, dtype=float)
        c2, e2 = S.initialize_pqeq(pq, chi_full[:pq.ntype + 1], eta_full[:pq.ntype + 1], rctap, CTap)
        chi_full[:pq.ntype + 1], eta_full[:pq.ntype + 1] = c2, e2
        chi, eta = chi_full, eta_full
    npt = np.zeros(ff.nso + 1, dtype=np.int64)
    for t in types0:
        npt[t] += 1
    rc, rc2, maxrc = S.cutoff_length(ff, npt)
    tables = S.potential_table(ff, rctap, CTap)
    cutoff_vpar30 = S.CUTOF2_BO * ff.vpar30
    pff = PackedFF(ff, rc2, tables, rctap, cutoff_vpar30, pqeq=pq, chi=chi, eta=eta)
    nprocs = int(np.prod(vprocs))
    boxes = [PackedBox(lattice, vprocs, r, maxrc, rctap) for r in range(nprocs)]
    ranks = []
    for r in range(nprocs):
        g = gen["ranks"][r]
        if g is None:
            ranks.append(None)
            continue
        b = boxes[r]
        obox = np.array(list(b.struct.OBOX))
        rn = g["pos_local"] + obox                       # xs2xu, src/main.F90:637-654
        H = b.H
        pos = np.empty((3, len(rn)))
        for c in range(3):
            pos[c] = (H[c, 0] * rn[:, 0] + H[c, 1] * rn[:, 1]) + H[c, 2] * rn[:, 2]
        n = len(rn)
        ranks.append(dict(atype=g["atype"].copy(), pos=np.ascontiguousarray(pos), v=np.zeros((3, n)), q=np.zeros(n)))
    return System(ff=ff, pff=pff, boxes=boxes, vprocs=tuple(vprocs), lattice=lattice, ranks=ranks,
                  natoms=gen["natoms"], maxrc=maxrc, rctap=rctap, mass=ff.mass, pqeq=pq)


def reference_path(*p):
    return os.path.join("/root/reference", *p)
