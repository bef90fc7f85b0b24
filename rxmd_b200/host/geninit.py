"""Synthetic-geometry generator: restatement of the reference's `geninit` tool.

Reference: init/geninit.F90:399-575 (read .xyz in normalised coordinates, replicate by
`-mc`, shift the minimum to 0, wrap, add 1e-9, split by `-vprocs`) and the rxff.bin
stream layout (init/geninit.F90:541-565, reader src/fileio.F90:477-536).

geninit cannot be compiled in this environment (no Fortran), so the harness creates
the very same per-rank atom sets directly in memory; `write_rxff_bin`/`read_rxff_bin`
keep the on-disk format available for hosts that want the file.
"""
from __future__ import annotations

import struct
import numpy as np

from .setup import get_box_params


def read_xyz(path, atom_names, real_coords=False):
    """Returns (types[int], pos0[n,3] normalised, (L1,L2,L3,al,be,ga)).

    `atom_names[1..nso]` is the ffield element order (getAtomNames, init/geninit.F90:207-241).
    `real_coords` handles the reference's real-coordinate files (ice-1h.xyz, rdx.xyz), which
    geninit would first pass through `-n` (convertAndDumpCoordinate).
    """
    with open(path) as fh:
        first = fh.readline().strip()
        natoms = int(first.split()[0])
        lat = [float(x) for x in fh.readline().split()[:6]]
        types = np.zeros(natoms, dtype=np.int32)
        pos = np.zeros((natoms, 3))
        for i in range(natoms):
            tok = fh.readline().split()
            name = tok[0]
            for j in range(1, len(atom_names)):
                if name == atom_names[j]:
                    types[i] = j
                    break
            else:
                raise ValueError(f"element {name} not in ffield")
            pos[i] = [float(tok[1]), float(tok[2]), float(tok[3])]
    if real_coords:
        H = get_box_params(*lat)
        Hi = np.linalg.inv(H)
        pos = pos @ Hi.T
    return types, pos, tuple(lat)


def replicate(types0, pos0, lat, mc, vprocs, no_shift=False, displace=None):
    """init/geninit.F90:446-527.  Returns dict with per-rank arrays.

    out['ranks'][r] = dict(pos_local[n,3] (normalised, minus OBOX), atype[n] (double,
    type + gid*1e-13 + 1e-14)), and the replicated lattice constants.
    """
    n0 = len(types0)
    mc = np.asarray(mc, dtype=np.int64)
    mctot = int(mc.prod())
    ix, iy, iz, ia = np.meshgrid(np.arange(mc[0]), np.arange(mc[1]), np.arange(mc[2]),
                                 np.arange(n0), indexing="ij")
    ix, iy, iz, ia = ix.ravel(), iy.ravel(), iz.ravel(), ia.ravel()
    ntot = n0 * mctot
    pos1 = np.empty((ntot, 3))
    pos1[:, 0] = (pos0[ia, 0] + ix) / mc[0]
    pos1[:, 1] = (pos0[ia, 1] + iy) / mc[1]
    pos1[:, 2] = (pos0[ia, 2] + iz) / mc[2]
    gid = np.arange(1, ntot + 1, dtype=np.float64)
    atype = types0[ia].astype(np.float64) + gid * 1e-13 + 1e-14
    if not no_shift:
        pos1 -= pos1.min(axis=0)
    pos1 = np.fmod(pos1, 1.0) + 1e-9
    if displace is not None:
        # synthetic thermal disorder (not part of geninit): `displace` = [ntot,3] normalised offsets; wrap back into [0,1)
        pos1 = pos1 + displace(ntot)
        pos1 = pos1 - np.floor(pos1)
        pos1[pos1 >= 1.0] = 0.0
    vp = np.asarray(vprocs, dtype=np.int64)
    cell = (pos1 * vp).astype(np.int64)
    sid = cell[:, 0] + cell[:, 1] * vp[0] + cell[:, 2] * vp[0] * vp[1]
    lbox = 1.0 / vp
    ranks = []
    for r in range(int(vp.prod())):
        sel = np.nonzero(sid == r)[0]
        obox = lbox * cell[sel[0]] if len(sel) else np.zeros(3)
        ranks.append(dict(pos_local=pos1[sel] - obox, atype=atype[sel].copy()))
    L = (lat[0] * mc[0], lat[1] * mc[1], lat[2] * mc[2], lat[3], lat[4], lat[5])
    return dict(ranks=ranks, lattice=L, natoms=ntot)


def write_rxff_bin(path, gen, vprocs, current_step=0):
    """rxff.bin stream (init/geninit.F90:541-565); 64-bit safe unlike the reference."""
    nprocs = len(gen["ranks"])
    with open(path, "wb") as fh:
        fh.write(struct.pack("<4i", nprocs, *[int(v) for v in vprocs]))
        fh.write(struct.pack(f"<{nprocs}i", *[len(r["atype"]) for r in gen["ranks"]]))
        fh.write(struct.pack("<i", current_step))
        fh.write(struct.pack("<6d", *gen["lattice"]))
        for r in gen["ranks"]:
            n = len(r["atype"])
            rec = np.zeros((n, 10))
            rec[:, 0:3] = r["pos_local"]
            if "v" in r:
                rec[:, 3:6] = r["v"]
            rec[:, 6] = r.get("q", 0.0)
            rec[:, 7] = r["atype"]
            fh.write(rec.astype("<f8").tobytes())


def read_rxff_bin(path):
    """ReadBIN's file layout (src/fileio.F90:477-536)."""
    with open(path, "rb") as fh:
        nprocs, vx, vy, vz = struct.unpack("<4i", fh.read(16))
        nat = struct.unpack(f"<{nprocs}i", fh.read(4 * nprocs))
        (step,) = struct.unpack("<i", fh.read(4))
        lat = struct.unpack("<6d", fh.read(48))
        ranks = []
        for n in nat:
            rec = np.frombuffer(fh.read(80 * n), dtype="<f8").reshape(n, 10)
            ranks.append(dict(pos_local=rec[:, 0:3].copy(), v=rec[:, 3:6].copy(), q=rec[:, 6].copy(),
                              atype=rec[:, 7].copy(), qsfp=rec[:, 8].copy(), qsfv=rec[:, 9].copy()))
    return dict(ranks=ranks, lattice=lat, vprocs=(vx, vy, vz), current_step=step,
                natoms=int(sum(nat)))
