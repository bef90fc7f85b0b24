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


def hash_normal(seed, gid, ncomp=3):
    """Standard normal deviates indexed by (seed, global atom id, component): a counter-based generator (splitmix64 mixing +
    Box-Muller), so that a rank can draw the displacements of ITS atoms without generating anybody else's -- the synthetic
    thermal disorder of the bench workloads must not depend on the decomposition."""
    gid = np.asarray(gid, dtype=np.uint64)

    def mix(x):
        x = (x + np.uint64(0x9E3779B97F4A7C15)) & np.uint64(0xFFFFFFFFFFFFFFFF)
        x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return x ^ (x >> np.uint64(31))

    out = np.empty((len(gid), ncomp))
    with np.errstate(over="ignore"):
        base = mix(np.uint64(seed) * np.uint64(0x2545F4914F6CDD1D) + gid * np.uint64(2 * ncomp))
        for c in range(ncomp):
            h1 = mix(base + np.uint64(2 * c))
            h2 = mix(base + np.uint64(2 * c + 1))
            u1 = ((h1 >> np.uint64(11)).astype(np.float64) + 0.5) / 9007199254740992.0     # (0,1)
            u2 = ((h2 >> np.uint64(11)).astype(np.float64) + 0.5) / 9007199254740992.0
            out[:, c] = np.sqrt(-2.0 * np.log(u1)) * np.cos(2.0 * np.pi * u2)
    return out


def _axis_candidates(mc_k, lo, hi):
    """Unit-cell indices along one axis whose atoms can land in the normalised slab [lo, hi): the slab's own cells plus one
    cell of margin on either side (the shift to the minimum and the synthetic displacements move an atom by far less than a
    cell), periodic."""
    i0 = int(np.floor(lo * mc_k)) - 1
    i1 = int(np.ceil(hi * mc_k)) + 1
    if i1 - i0 >= mc_k:
        return np.arange(mc_k, dtype=np.int64)
    return np.unique(np.arange(i0, i1 + 1, dtype=np.int64) % mc_k)


def replicate(types0, pos0, lat, mc, vprocs, no_shift=False, displace=None, only_rank=None):
    """init/geninit.F90:446-527.  Returns dict with per-rank arrays.

    out['ranks'][r] = dict(pos_local[n,3] (normalised, minus OBOX), atype[n] (double,
    type + gid*1e-13 + 1e-14)), and the replicated lattice constants.

    Every rank's set is generated from the unit cells that can reach its sub-domain only (geninit itself loops over all
    atoms once and writes each to its rank, :493-527; here a rank never materialises the other ranks' atoms), so host
    memory and time scale with the atoms of the ranks asked for: `only_rank=r` builds rank r alone (the others are None).
    `displace(gid) -> [n,3]` gives normalised synthetic offsets per global atom id (not part of geninit).
    """
    n0 = len(types0)
    mc = np.asarray(mc, dtype=np.int64)
    mctot = int(mc.prod())
    ntot = n0 * mctot
    vp = np.asarray(vprocs, dtype=np.int64)
    nprocs = int(vp.prod())
    lbox = 1.0 / vp
    # the shift of the global minimum to 0 (:481-485): the minimum over all replicas is the unit cell's own, in replica 0
    shift = np.zeros(3) if no_shift else np.array([(pos0[:, k] + 0.0).min() / mc[k] for k in range(3)])
    ia0 = np.arange(n0, dtype=np.int64)
    ranks = [None] * nprocs
    for r in (range(nprocs) if only_rank is None else [int(only_rank)]):
        vid = np.array([r % vp[0], (r // vp[0]) % vp[1], r // (vp[0] * vp[1])], dtype=np.int64)
        cand = [_axis_candidates(int(mc[k]), vid[k] * lbox[k], (vid[k] + 1) * lbox[k]) if nprocs > 1 else np.arange(mc[k], dtype=np.int64)
                for k in range(3)]
        ix, iy, iz, ia = np.meshgrid(cand[0], cand[1], cand[2], ia0, indexing="ij")
        ix, iy, iz, ia = ix.ravel(), iy.ravel(), iz.ravel(), ia.ravel()
        pos1 = np.empty((len(ia), 3))
        pos1[:, 0] = (pos0[ia, 0] + ix) / mc[0]
        pos1[:, 1] = (pos0[ia, 1] + iy) / mc[1]
        pos1[:, 2] = (pos0[ia, 2] + iz) / mc[2]
        gidi = ((ix * mc[1] + iy) * mc[2] + iz) * n0 + ia + 1          # geninit's loop order: atom index fastest (:446-460)
        pos1 -= shift
        pos1 = np.fmod(pos1, 1.0) + 1e-9
        if displace is not None:
            # synthetic thermal disorder (not part of geninit); wrap back into [0,1)
            pos1 = pos1 + displace(gidi)
            pos1 = pos1 - np.floor(pos1)
            pos1[pos1 >= 1.0] = 0.0
        cell = (pos1 * vp).astype(np.int64)
        sid = cell[:, 0] + cell[:, 1] * vp[0] + cell[:, 2] * vp[0] * vp[1]
        sel = np.nonzero(sid == r)[0]
        sel = sel[np.argsort(gidi[sel], kind="stable")]                  # rank-local order = global atom order, like geninit
        gid = gidi[sel].astype(np.float64)
        atype = types0[ia[sel]].astype(np.float64) + gid * 1e-13 + 1e-14
        obox = lbox * vid
        ranks[r] = dict(pos_local=pos1[sel] - obox, atype=atype)
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


def read_rxff_bin(path, only_rank=None):
    """ReadBIN's file layout (src/fileio.F90:477-536).  `only_rank=r` reads rank r's records alone, like each MPI rank of the
    reference does (:499-505, an MPI_Scan of the per-rank byte counts), but with 64-bit offsets (the reference's default
    integers overflow beyond 2^31 B = 26.8 M atoms); the other ranks' entries are None."""
    with open(path, "rb") as fh:
        nprocs, vx, vy, vz = struct.unpack("<4i", fh.read(16))
        nat = struct.unpack(f"<{nprocs}i", fh.read(4 * nprocs))
        (step,) = struct.unpack("<i", fh.read(4))
        lat = struct.unpack("<6d", fh.read(48))
        data0 = fh.tell()
        ranks = [None] * nprocs
        for r in (range(nprocs) if only_rank is None else [int(only_rank)]):
            fh.seek(data0 + 80 * int(sum(int(x) for x in nat[:r])))
            n = nat[r]
            rec = np.frombuffer(fh.read(80 * n), dtype="<f8").reshape(n, 10)
            ranks[r] = dict(pos_local=rec[:, 0:3].copy(), v=rec[:, 3:6].copy(), q=rec[:, 6].copy(),
                            atype=rec[:, 7].copy(), qsfp=rec[:, 8].copy(), qsfv=rec[:, 9].copy())
    return dict(ranks=ranks, lattice=lat, vprocs=(vx, vy, vz), current_step=step,
                natoms=int(sum(nat)), natoms_per_rank=tuple(int(x) for x in nat))
