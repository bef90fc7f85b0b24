"""Host-side one-time setup: the restatement of INITSYSTEM's numeric parts.

In production these run once in the Fortran host (reference src/init.F90); the
hot-path library only consumes their results through `rxg_set_forcefield` /
`rxg_set_box`.  Restated here so the harness can produce the very same inputs:

  taper coefficients      src/init.F90:28-38
  rank topology           src/init.F90:74-100
  CUTOFFLENGTH            src/init.F90:363-418
  POTENTIALTABLE          src/init.F90:421-522
  GetNonbondingMesh       src/init.F90:525-607
  GetBoxParams            src/init.F90:610-633
  UpdateBoxParams/matinv  src/init.F90:636-668, src/main.F90:557-579
  initialize_pqeq tables  src/module.F90:488-613
"""
from __future__ import annotations

from dataclasses import dataclass, field
import math
import numpy as np

from .ffield import ForceField

# compile-time constants of the reference (src/module.F90:44-45,60-64,80-87,251,281-282,677-683)
MAXLAYERS = 5
MAXLAYERS_NB = 10
MINBOSIG = 1e-3
MINBO0 = 1e-4
CUTOF2_ESUB = 1e-4
CUTOF2_BO = 1e-3
NMINCELL = 4
NTABLE = 5000
RCTAP0 = 10.0
RCTAP0_PQEQ = 12.5
CCLMB0 = 332.0638
CCLMB0_QEQ = 14.4
CECHRGE = 23.02
UTIME = 1.0e3 / 20.455
UTEMP0 = 503.398008
UTEMP = UTEMP0 * 2.0 / 3.0
USTRS = 6.94728103
UDENS = 1.66053886
EEV_KCAL = 23.060538
LAMBDA_PQEQ = 0.462770


def get_box_params(la, lb, lc, a1, a2, a3):
    """GetBoxParams, src/init.F90:610-633. Returns H with H[i-1, j-1] = H(i,j)."""
    pi = math.atan(1.0) * 4.0
    lal, lbe, lga = a1 * pi / 180.0, a2 * pi / 180.0, a3 * pi / 180.0
    hh1 = lc * (math.cos(lal) - math.cos(lbe) * math.cos(lga)) / math.sin(lga)
    hh2 = lc * math.sqrt(1.0 - math.cos(lal) ** 2 - math.cos(lbe) ** 2 - math.cos(lga) ** 2
                         + 2 * math.cos(lal) * math.cos(lbe) * math.cos(lga)) / math.sin(lga)
    H = np.zeros((3, 3))
    H[0, 0] = la;                H[1, 0] = 0.0;               H[2, 0] = 0.0
    H[0, 1] = lb * math.cos(lga); H[1, 1] = lb * math.sin(lga); H[2, 1] = 0.0
    H[0, 2] = lc * math.cos(lbe); H[1, 2] = hh1;              H[2, 2] = hh2
    return H


def matinv(m1):
    """matinv, src/main.F90:557-579 (adjugate / determinant, same expression order)."""
    m = lambda i, j: m1[i - 1, j - 1]
    m2 = np.zeros((3, 3))
    m2[0, 0] = m(2, 2) * m(3, 3) - m(2, 3) * m(3, 2)
    m2[0, 1] = m(1, 3) * m(3, 2) - m(1, 2) * m(3, 3)
    m2[0, 2] = m(1, 2) * m(2, 3) - m(1, 3) * m(2, 2)
    m2[1, 0] = m(2, 3) * m(3, 1) - m(2, 1) * m(3, 3)
    m2[1, 1] = m(1, 1) * m(3, 3) - m(1, 3) * m(3, 1)
    m2[1, 2] = m(1, 3) * m(2, 1) - m(1, 1) * m(2, 3)
    m2[2, 0] = m(2, 1) * m(3, 2) - m(2, 2) * m(3, 1)
    m2[2, 1] = m(1, 2) * m(3, 1) - m(1, 1) * m(3, 2)
    m2[2, 2] = m(1, 1) * m(2, 2) - m(1, 2) * m(2, 1)
    detm = (m(1, 1) * m(2, 2) * m(3, 3) + m(1, 2) * m(2, 3) * m(3, 1)
            + m(1, 3) * m(2, 1) * m(3, 2) - m(1, 3) * m(2, 2) * m(3, 1)
            - m(1, 2) * m(2, 1) * m(3, 3) - m(1, 1) * m(2, 3) * m(3, 2))
    return m2 / detm


def cutoff_length(ff: ForceField, natoms_per_type):
    """CUTOFFLENGTH, src/init.F90:363-418. Returns rc[0..nboty] (1-based), rc2, maxrc."""
    rc = np.zeros(ff.nboty + 1)
    rc2 = np.zeros(ff.nboty + 1)
    for ity in range(1, ff.nso + 1):
        for jty in range(ity, ff.nso + 1):
            inxn = ff.inxn2[ity, jty]
            if inxn == 0:
                continue
            dr = 1.0
            BOsig = 1.0
            while BOsig > MINBOSIG:
                dr = dr + 0.01
                BOsig = math.exp(ff.pbo1[inxn] * (dr / ff.r0s[ity, jty]) ** ff.pbo2[inxn])
            rc[inxn] = dr
            rc2[inxn] = dr * dr
    for ity in range(1, ff.nso + 1):
        if natoms_per_type[ity] == 0:
            for jty in range(1, ff.nso + 1):
                inxn = ff.inxn2[ity, jty]
                if inxn != 0:
                    rc[inxn] = 0.0
                inxn = ff.inxn2[jty, ity]
                if inxn != 0:
                    rc[inxn] = 0.0
    return rc, rc2, float(rc.max())


def taper(rctap):
    """CTap(0:7), src/init.F90:36-38."""
    return np.array([1.0, 0.0, 0.0, 0.0, -35.0 / rctap ** 4, 84.0 / rctap ** 5,
                     -70.0 / rctap ** 6, 20.0 / rctap ** 7])


def potential_table(ff: ForceField, rctap, CTap):
    """POTENTIALTABLE, src/init.F90:421-522.

    Returns (TBL_Evdw, TBL_Eclmb, TBL_Eclmb_QEq, UDR, UDRi) with
    TBL_Evdw[c, i, inxn] == TBL_Evdw(c,i,inxn) for i in 1..NTABLE (index 0 unused).
    """
    rctap2 = rctap ** 2
    UDR = rctap2 / NTABLE
    UDRi = 1.0 / UDR
    nb = ff.nboty
    T_vdw = np.zeros((2, NTABLE + 1, nb + 1))
    T_clmb = np.zeros((2, NTABLE + 1, nb + 1))
    T_qeq = np.zeros((NTABLE + 1, nb + 1))
    i = np.arange(1, NTABLE + 1, dtype=np.float64)
    dr2 = UDR * i
    dr1 = np.sqrt(dr2)
    dr3 = dr1 * dr2
    dr4 = dr2 * dr2
    dr5 = dr1 * dr2 * dr2
    dr6 = dr2 * dr2 * dr2
    dr7 = dr1 * dr2 * dr2 * dr2
    Tap = CTap[7] * dr7 + CTap[6] * dr6 + CTap[5] * dr5 + CTap[4] * dr4 + CTap[0]
    dTap = 7.0 * CTap[7] * dr5 + 6.0 * CTap[6] * dr4 + 5.0 * CTap[5] * dr3 + 4.0 * CTap[4] * dr2
    for ity in range(1, ff.nso + 1):
        for jty in range(ity, ff.nso + 1):
            inxn = ff.inxn2[ity, jty]
            if inxn == 0:
                continue
            gamWij = ff.gamW[ity, jty]; alphaij = ff.alpij[ity, jty]
            Dij0 = ff.Dij[ity, jty]; rvdW0 = ff.rvdW[ity, jty]
            gamwinvp = (1.0 / gamWij) ** ff.pvdW1
            rij_vd1 = dr2 ** ff.pvdW1h
            fn13 = (rij_vd1 + gamwinvp) ** ff.pvdW1inv
            exp1 = np.exp(alphaij * (1.0 - fn13 / rvdW0))
            exp2 = np.sqrt(exp1)
            dr3gamij = (dr3 + ff.gamij[ity, jty]) ** (-1.0 / 3.0)
            T_vdw[0, 1:, inxn] = Tap * Dij0 * (exp1 - 2.0 * exp2)
            T_clmb[0, 1:, inxn] = Tap * CCLMB0 * dr3gamij
            T_qeq[1:, inxn] = Tap * CCLMB0_QEQ * dr3gamij
            dfn13 = ((rij_vd1 + gamwinvp) ** (ff.pvdW1inv - 1.0)) * (dr2 ** (ff.pvdW1h - 1.0))
            T_vdw[1, 1:, inxn] = Dij0 * (dTap * (exp1 - 2.0 * exp2)
                                         - Tap * (alphaij / rvdW0) * (exp1 - exp2) * dfn13)
            T_clmb[1, 1:, inxn] = CCLMB0 * dr3gamij * (dTap - (dr3gamij ** 3) * Tap * dr1)
            if ff.isLG:
                if ity > 4 or jty > 4:
                    continue
                dr_lg = 2 * math.sqrt(ff.Re_lg[ity] * ff.Re_lg[jty])
                dr6_lg = dr_lg ** 6
                Elg = -ff.C_lg[ity, jty] / (dr6 + dr6_lg)
                E_core = ff.ecore[ity, jty] * np.exp(ff.acore[ity, jty] * (1.0 - (dr1 / ff.rcore[ity, jty])))
                dElg = ff.C_lg[ity, jty] * (6.0 * dr5) / (dr6 + dr6_lg) ** 2 / dr1
                dE_core = -ff.acore[ity, jty] * E_core / ff.rcore[ity, jty] / dr1
                T_vdw[0, 1:, inxn] += Tap * (Elg + E_core)
                T_vdw[1, 1:, inxn] += dTap * Elg + Tap * dElg + dTap * E_core + Tap * dE_core
    return T_vdw, T_clmb, T_qeq, UDR, UDRi


def nonbonding_mesh(lata, latb, latc, vprocs, rctap):
    """GetNonbondingMesh, src/init.F90:525-607. Returns nbcc, nblcsize (normalised), nbmesh[n,3]."""
    nblcsize = np.array([3.0, 3.0, 3.0])
    lpn = np.array([lata / vprocs[0], latb / vprocs[1], latc / vprocs[2]])
    nbcc = (lpn / nblcsize).astype(np.int64)
    nblcsize = lpn / nbcc
    imesh = (rctap / nblcsize).astype(np.int64) + 1
    mesh = []
    for i in range(-imesh[0], imesh[0] + 1):
        for j in range(-imesh[1], imesh[1] + 1):
            for k in range(-imesh[2], imesh[2] + 1):
                ii = [i, j, k]
                for a in range(3):
                    if ii[a] > 0:
                        ii[a] -= 1
                    elif ii[a] < 0:
                        ii[a] += 1
                rr = np.array(ii, dtype=np.float64) * nblcsize
                d2 = (rr[0] * rr[0] + rr[1] * rr[1]) + rr[2] * rr[2]
                if d2 <= rctap ** 2:
                    mesh.append((i, j, k))
    nbmesh = np.array(mesh, dtype=np.int32).reshape(-1, 3)
    nblcsize = nblcsize / np.array([lata, latb, latc])
    return nbcc.astype(np.int32), nblcsize, nbmesh


def rank_topology(myid, vprocs):
    """vID / myparity / target_node, src/init.F90:74-100."""
    vx, vy, vz = vprocs
    vID = [myid % vx, (myid // vx) % vy, myid // (vx * vy)]
    parity = [v % 2 for v in vID]
    target = []
    for i in range(3):
        for j in (1, -1):
            l = list(vID)
            l[i] = (vID[i] + j + vprocs[i]) % vprocs[i]
            target.append(l[0] + l[1] * vx + l[2] * vx * vy)
    return vID, parity, target


@dataclass
class PQEqParams:
    """`module pqeq_vars` after get_pqeq_parms + initialize_pqeq (src/cmdline.F90:168-236, src/module.F90:448-613)."""
    ntype: int = 0
    elem: list = field(default_factory=list)
    polarizable: np.ndarray = None
    X0: np.ndarray = None
    J0: np.ndarray = None
    Z: np.ndarray = None
    Rc: np.ndarray = None
    Rs: np.ndarray = None
    Ks: np.ndarray = None
    alphacc: np.ndarray = None
    alphasc: np.ndarray = None
    alphass: np.ndarray = None
    inxnpqeq: np.ndarray = None
    T_pcc: np.ndarray = None
    T_psc: np.ndarray = None
    T_pss: np.ndarray = None


def read_pqeq_parms(path):
    """get_pqeq_parms, src/cmdline.F90:168-236 (every listed element is flagged polarizable, :212)."""
    p = PQEqParams()
    rows = []
    nparms = None
    with open(path) as fh:
        for ln in fh:
            s = ln.strip()
            if not s or s.startswith("#"):
                continue
            if "NPARMS" in s:
                nparms = int(s.split()[1])
                continue
            if nparms is not None and len(rows) < nparms:
                rows.append(s.split())
    n = nparms
    p.ntype = n
    p.elem = [""] + [r[0] for r in rows]
    arr = lambda k: np.array([0.0] + [float(r[k]) for r in rows])
    p.polarizable = np.array([False] + [True] * n)
    p.X0, p.J0, p.Z, p.Rc, p.Rs, p.Ks = arr(2), arr(3), arr(4), arr(5), arr(6), arr(7)
    return p


def truncate_pqeq(p: PQEqParams, n):
    """Keep the first n element rows (see system.build_system)."""
    p.ntype = n
    p.elem = p.elem[:n + 1]
    for k in ("polarizable", "X0", "J0", "Z", "Rc", "Rs", "Ks"):
        setattr(p, k, getattr(p, k)[:n + 1])


def initialize_pqeq(p: PQEqParams, chi, eta, rctap, CTap):
    """set_alphaij_pqeq + initialize_pqeq, src/module.F90:448-613.

    `chi`/`eta` are the (nso+1)-long ffield arrays (eta already doubled by GETPARAMS);
    returns updated copies.  Note the reference doubles eta a second time (:523).
    """
    from scipy.special import erf
    n = p.ntype
    acc = np.zeros((n + 1, n + 1)); asc = np.zeros((n + 1, n + 1)); ass = np.zeros((n + 1, n + 1))
    for i in range(1, n + 1):
        aci = 0.5 * LAMBDA_PQEQ / p.Rc[i] ** 2
        asi = 0.5 * LAMBDA_PQEQ / p.Rs[i] ** 2
        for j in range(1, n + 1):
            acj = 0.5 * LAMBDA_PQEQ / p.Rc[j] ** 2
            asj = 0.5 * LAMBDA_PQEQ / p.Rs[j] ** 2
            acc[i, j] = math.sqrt((aci * acj) / (aci + acj))
            if p.polarizable[i] and p.polarizable[j]:
                ass[i, j] = math.sqrt((asi * asj) / (asi + asj))
            if p.polarizable[i]:
                asc[i, j] = math.sqrt((asi * acj) / (asi + acj))
    p.alphacc, p.alphasc, p.alphass = acc, asc, ass
    chi = chi.copy(); eta = eta.copy()
    for i in range(1, n + 1):
        if not p.polarizable[i]:
            p.Z[i] = 0.0; p.Ks[i] = 0.0
        else:
            chi[i] = p.X0[i]; eta[i] = p.J0[i]
    eta = 2.0 * eta
    inx = np.zeros((n + 1, n + 1), dtype=np.int32)
    c = 0
    for i in range(1, n + 1):
        for j in range(i, n + 1):
            c += 1
            inx[i, j] = c; inx[j, i] = c
    p.inxnpqeq = inx
    n2 = n * n
    rctap2 = rctap ** 2
    UDR = rctap2 / NTABLE
    T = [np.zeros((n2 + 1, NTABLE + 1, 2)) for _ in range(3)]
    ii = np.arange(1, NTABLE + 1, dtype=np.float64)
    dr2 = UDR * ii; dr1 = np.sqrt(dr2)
    dr3 = dr1 * dr2; dr4 = dr2 * dr2; dr5 = dr1 * dr2 * dr2; dr6 = dr2 * dr2 * dr2; dr7 = dr1 * dr2 * dr2 * dr2
    Tap = CTap[7] * dr7 + CTap[6] * dr6 + CTap[5] * dr5 + CTap[4] * dr4 + CTap[0]
    dTap = 7.0 * CTap[7] * dr5 + 6.0 * CTap[6] * dr4 + 5.0 * CTap[5] * dr3 + 4.0 * CTap[4] * dr2
    dr1i = 1.0 / dr1
    clmb = dr1i
    dclmb = -dr1i * dr1i * dr1i
    sqrtpi_inv = 1.0 / math.sqrt(3.14159265358979)
    for i in range(1, n + 1):
        for j in range(i, n + 1):
            k = inx[i, j]
            for t, A in zip(T, (acc[i, j], asc[i, j], ass[i, j])):
                screen = erf(A * dr1)
                dscreen = 2.0 * A * sqrtpi_inv * np.exp(-A * A * dr2) * dr1i
                t[k, 1:, 0] = clmb * screen * Tap
                t[k, 1:, 1] = dclmb * screen * Tap + clmb * dscreen * Tap + clmb * screen * dTap
    p.T_pcc, p.T_psc, p.T_pss = T
    return chi, eta
