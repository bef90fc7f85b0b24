"""ReaxFF `ffield` reader: the host-side restatement of GETPARAMS.

The Fortran host keeps this job in production (reference src/param.F90:2-375); this
module exists so that the Python harness can feed the C-ABI library and the oracle
without a Fortran toolchain.  Every array is laid out exactly as the Fortran module
`parameters` holds it (1-based, column-major), so the buffers can be handed to
`rxg_set_forcefield` unchanged -- the same pointers a Fortran shim would pass.

Fixed-column semantics follow the Fortran edit descriptors (src/param.F90:344-351):
a blank field reads as 0, a field without a decimal point is scaled by 10**-d.
"""
from __future__ import annotations

from dataclasses import dataclass, field
import numpy as np

PI_RXMD = 3.14159265358979  # src/module.F90:90 (note: truncated pi, on purpose)


def _ffield(text: str, start: int, width: int, decimals: int) -> float:
    """Read one Fortran Fw.d field starting at 0-based column `start`."""
    s = text[start:start + width]
    s = s.strip()
    if not s:
        return 0.0
    s = s.replace("d", "e").replace("D", "e")
    if "." in s or "e" in s.lower():
        return float(s)
    return float(int(s)) / (10.0 ** decimals)


def _ifield(text: str, start: int, width: int) -> int:
    s = text[start:start + width].strip()
    return int(s) if s else 0


def _floats(text: str, skip: int, n: int, width: int = 9, decimals: int = 4):
    return [_ffield(text, skip + k * width, width, decimals) for k in range(n)]


@dataclass
class ForceField:
    """Flattened `module parameters` (src/module.F90:620-723), Fortran layouts."""
    header: str = ""
    nso: int = 0
    nboty: int = 0
    nvaty: int = 0
    ntoty: int = 0
    nhbty: int = 0
    isLG: bool = False
    atmname: list = field(default_factory=list)
    vpar: np.ndarray = None
    # everything else is attached dynamically as numpy arrays (see read_ffield)

    def F(self, name):
        return getattr(self, name)


def read_ffield(path: str, isLG: bool = False) -> ForceField:
    """Restates GETPARAMS (reference src/param.F90:2-375) line by line."""
    with open(path, "r") as fh:
        lines = [ln.rstrip("\n") for ln in fh]
    it = iter(lines)
    ff = ForceField()
    ff.isLG = isLG
    ff.header = next(it)[:100]                               # param.F90:40
    npar = int(next(it).split()[0])                          # param.F90:42 (list-directed)
    vpar = np.zeros(npar + 1)
    for i0 in range(1, npar + 1):                            # param.F90:46-48, format 1300 (f10.4)
        vpar[i0] = _ffield(next(it), 0, 10, 4)
    ff.vpar = vpar
    ff.pvdW1 = vpar[29]; ff.pvdW1h = 0.5 * vpar[29]; ff.pvdW1inv = 1.0 / vpar[29]   # :51-53
    ff.vpar30 = vpar[30]                                     # :56
    nso = _ifield(next(it), 0, 3)                            # :59
    ff.nso = nso
    z1 = lambda: np.zeros(nso + 1)
    z2 = lambda: np.zeros((nso + 1, nso + 1))
    rat, rapt, vnq = z1(), z1(), z1()
    Val, Valboc, mass = z1(), z1(), z1()
    bo131, bo132, bo133 = z1(), z1(), z1()
    Vale, plp2 = z1(), z1()
    povun2, povun5 = z1(), z1()
    pval3, pval5, Valval = z1(), z1(), z1()
    rvdw1, eps, alf, vop = z1(), z1(), z1(), z1()
    chi, eta, gam = z1(), z1(), z1()
    C_lg = z2(); Re_lg = z1(); rcore2, ecore2, acore2 = z1(), z1(), z1()
    plp1 = np.full(nso + 1, vpar[16]); povun3 = np.full(nso + 1, vpar[33])      # :90-95
    povun4 = np.full(nso + 1, vpar[32]); povun6 = np.full(nso + 1, vpar[7])
    povun7 = np.full(nso + 1, vpar[9]); povun8 = np.full(nso + 1, vpar[10])
    next(it); next(it); next(it)                             # :98-100
    atmname = [""] * (nso + 1)
    for i1 in range(1, nso + 1):                             # :102-114
        ln = next(it)
        atmname[i1] = ln[1:3].strip()
        (rat[i1], Val[i1], mass[i1], rvdw1[i1], eps[i1], gam[i1], rapt[i1], Vale[i1]) = _floats(ln, 3, 8)
        ln = next(it)
        a = _floats(ln, 3, 8)
        alf[i1], vop[i1], Valboc[i1], povun5[i1], chi[i1], eta[i1] = a[0], a[1], a[2], a[3], a[5], a[6]
        ln = next(it)
        a = _floats(ln, 3, 8)
        vnq[i1], plp2[i1], bo131[i1], bo132[i1], bo133[i1] = a[0], a[1], a[3], a[4], a[5]
        ln = next(it)
        a = _floats(ln, 3, 8)
        povun2[i1], pval3[i1], Valval[i1], pval5[i1] = a[0], a[1], a[3], a[4]
        if isLG:
            rcore2[i1], ecore2[i1], acore2[i1] = a[5], a[6], a[7]
            a = _floats(next(it), 3, 2)
            C_lg[i1, i1], Re_lg[i1] = a[0], a[1]
    ff.atmname = atmname
    for i1 in range(1, nso + 1):                             # :117-119
        if mass[i1] < 21.0 and Valboc[i1] != Valval[i1]:
            Valboc[i1] = Valval[i1]
    nlpopt = 0.5 * (Vale - Val)                              # :121
    Valangle = Valboc.copy()                                 # :123
    r0s, r0p, r0pp = z2(), z2(), z2()
    rvdW, Dij, alpij, gamW, gamij = z2(), z2(), z2(), z2(), z2()
    rcore, ecore, acore = z2(), z2(), z2()
    for i1 in range(1, nso + 1):                             # :126-148
        for i2 in range(1, nso + 1):
            r0s[i1, i2] = 0.5 * (rat[i1] + rat[i2])
            r0p[i1, i2] = 0.5 * (rapt[i1] + rapt[i2])
            r0pp[i1, i2] = 0.5 * (vnq[i1] + vnq[i2])
            rvdW[i1, i2] = np.sqrt(4.0 * rvdw1[i1] * rvdw1[i2])
            Dij[i1, i2] = np.sqrt(eps[i1] * eps[i2])
            alpij[i1, i2] = np.sqrt(alf[i1] * alf[i2])
            gamW[i1, i2] = np.sqrt(vop[i1] * vop[i2])
            gamij[i1, i2] = (gam[i1] * gam[i2]) ** (-1.5)
            if isLG:
                rcore[i1, i2] = np.sqrt(rcore2[i1] * rcore2[i2])
                ecore[i1, i2] = np.sqrt(ecore2[i1] * ecore2[i2])
                acore[i1, i2] = np.sqrt(acore2[i1] * acore2[i2])
    nboty = _ifield(next(it), 0, 3)                          # :151
    ff.nboty = nboty
    b1 = lambda: np.zeros(nboty + 1)
    pbo1, pbo2, pbo3, pbo4, pbo5, pbo6, bom = b1(), b1(), b1(), b1(), b1(), b1(), b1()
    pboc3, pboc4, pboc5 = b1(), b1(), b1()
    Desig, Depi, Depipi, pbe1, pbe2 = b1(), b1(), b1(), b1(), b1()
    povun1, ovc, v13cor = b1(), b1(), b1()
    next(it)                                                 # :160
    inxn2 = np.zeros((nso + 1, nso + 1), dtype=np.int32)
    for ih in range(1, nboty + 1):                           # :164-170, formats 1400/1450
        ln = next(it)
        typea, typeb = _ifield(ln, 0, 3), _ifield(ln, 3, 3)
        (Desig[ih], Depi[ih], Depipi[ih], pbe1[ih], pbo5[ih], v13cor[ih], pbo6[ih], povun1[ih]) = _floats(ln, 6, 8)
        ln = next(it)
        a = _floats(ln, 6, 8)
        pbe2[ih], pbo3[ih], pbo4[ih], bom[ih], pbo1[ih], pbo2[ih], ovc[ih] = a[:7]
        inxn2[typea, typeb] = ih
        inxn2[typeb, typea] = ih
    pboc1 = np.full(nboty + 1, vpar[1]); pboc2 = np.full(nboty + 1, vpar[2])    # :174-175
    ff.vpar1, ff.vpar2 = vpar[1], vpar[2]                    # :178-179
    for i1 in range(1, nso + 1):                             # :181-190
        for i2 in range(1, nso + 1):
            inxn = inxn2[i1, i2]
            if inxn != 0:
                pboc3[inxn] = np.sqrt(bo132[i1] * bo132[i2])
                pboc4[inxn] = np.sqrt(bo131[i1] * bo131[i2])
                pboc5[inxn] = np.sqrt(bo133[i1] * bo133[i2])
    nodmty = _ifield(next(it), 0, 3)                         # :194
    for _ in range(nodmty):                                  # :195-217
        ln = next(it)
        n1, n2 = _ifield(ln, 0, 3), _ifield(ln, 3, 3)
        a = _floats(ln, 6, 7 if isLG else 6)
        deodmh, rodmh, godmh, rsig, rpi, rpi2 = a[:6]
        if isLG:
            C_lg[n1, n2] = a[6]; C_lg[n2, n1] = a[6]
        if rsig > 0.0: r0s[n1, n2] = rsig; r0s[n2, n1] = rsig
        if rpi > 0.0: r0p[n1, n2] = rpi; r0p[n2, n1] = rpi
        if rpi2 > 0.0: r0pp[n1, n2] = rpi2; r0pp[n2, n1] = rpi2
        if rodmh > 0.0: rvdW[n1, n2] = 2.0 * rodmh; rvdW[n2, n1] = 2.0 * rodmh
        if deodmh > 0.0: Dij[n1, n2] = deodmh; Dij[n2, n1] = deodmh
        if godmh > 0.0: alpij[n1, n2] = godmh; alpij[n2, n1] = godmh
    cBOp1, cBOp3, cBOp5 = b1(), b1(), b1()
    pbo2h, pbo4h, pbo6h = b1(), b1(), b1()
    switch = np.zeros((4, nboty + 1))                        # switch(1:3,inxn) -> [c, inxn]
    for i in range(1, nso + 1):                              # :227-261
        for j in range(1, nso + 1):
            inxn = inxn2[i, j]
            if inxn == 0:
                continue
            if rat[i] > 0.0 and rat[j] > 0.0: switch[1, inxn] = 1
            if rapt[i] > 0.0 and rapt[j] > 0.0: switch[2, inxn] = 1
            if vnq[i] > 0.0 and vnq[j] > 0.0: switch[3, inxn] = 1
            cBOp1[inxn] = 0.0 if r0s[i, j] <= 0.0 else pbo1[inxn] / (r0s[i, j] ** pbo2[inxn])
            cBOp3[inxn] = 0.0 if r0p[i, j] <= 0.0 else pbo3[inxn] / (r0p[i, j] ** pbo4[inxn])
            cBOp5[inxn] = 0.0 if r0pp[i, j] <= 0.0 else pbo5[inxn] / (r0pp[i, j] ** pbo6[inxn])
            pbo2h[inxn] = 0.5 * pbo2[inxn]
            pbo4h[inxn] = 0.5 * pbo4[inxn]
            pbo6h[inxn] = 0.5 * pbo6[inxn]
    inxn3 = np.zeros((nso + 1,) * 3, dtype=np.int32)
    nvaty = _ifield(next(it), 0, 3)                          # :265
    ff.nvaty = nvaty
    v1 = lambda: np.zeros(nvaty + 1)
    theta00, pval1, pval2, pcoa1, pval7, ppen1, pval4 = v1(), v1(), v1(), v1(), v1(), v1(), v1()
    for i in range(1, nvaty + 1):                            # :273-277, format 1500
        ln = next(it)
        i1, i2, i3 = _ifield(ln, 0, 3), _ifield(ln, 3, 3), _ifield(ln, 6, 3)
        (theta00[i], pval1[i], pval2[i], pcoa1[i], pval7[i], ppen1[i], pval4[i]) = _floats(ln, 9, 7)
        inxn3[i1, i2, i3] = i
        inxn3[i3, i2, i1] = i
    pval6 = np.full(nvaty + 1, vpar[15]); pval8 = np.full(nvaty + 1, vpar[34])   # :280-291
    pval9 = np.full(nvaty + 1, vpar[17]); pval10 = np.full(nvaty + 1, vpar[18])
    ppen2 = np.full(nvaty + 1, vpar[20]); ppen3 = np.full(nvaty + 1, vpar[21])
    ppen4 = np.full(nvaty + 1, vpar[22])
    pcoa2 = np.full(nvaty + 1, vpar[3]); pcoa3 = np.full(nvaty + 1, vpar[39])
    pcoa4 = np.full(nvaty + 1, vpar[31])
    theta00 = (PI_RXMD / 180.0) * theta00                    # :293
    ntoty = _ifield(next(it), 0, 3)                          # :296
    ff.ntoty = ntoty
    t1 = lambda: np.zeros(ntoty + 1)
    V1, V2, V3, ptor1, pcot1 = t1(), t1(), t1(), t1(), t1()
    inxn4 = np.zeros((nso + 1,) * 4, dtype=np.int32)
    for i in range(1, ntoty + 1):                            # :301-321, format 1600
        ln = next(it)
        i1, i2, i3, i4 = (_ifield(ln, 3 * k, 3) for k in range(4))
        a = _floats(ln, 12, 7)
        V1[i], V2[i], V3[i], ptor1[i], pcot1[i] = a[:5]
        if i1 == 0:
            for j1 in range(1, nso + 1):
                for j4 in range(1, nso + 1):
                    if inxn4[j1, i2, i3, j4] == 0 and inxn4[j1, i3, i2, j4] == 0:
                        inxn4[j1, i2, i3, j4] = i
                        inxn4[j4, i2, i3, j1] = i
                        inxn4[j1, i3, i2, j4] = i
                        inxn4[j4, i3, i2, j1] = i
        else:
            inxn4[i1, i2, i3, i4] = i
            inxn4[i4, i2, i3, i1] = i
            inxn4[i1, i3, i2, i4] = i
            inxn4[i4, i3, i2, i1] = i
    ptor2 = np.full(ntoty + 1, vpar[24]); ptor3 = np.full(ntoty + 1, vpar[25])   # :324-327
    ptor4 = np.full(ntoty + 1, vpar[26]); pcot2 = np.full(ntoty + 1, vpar[28])
    inxn3hb = np.zeros((nso + 1,) * 3, dtype=np.int32)
    nhbty = _ifield(next(it), 0, 3)                          # :331
    ff.nhbty = nhbty
    h1 = lambda: np.zeros(nhbty + 1)
    r0hb, phb1, phb2, phb3 = h1(), h1(), h1(), h1()
    for i in range(1, nhbty + 1):                            # :334-337
        ln = next(it)
        i1, i2, i3 = _ifield(ln, 0, 3), _ifield(ln, 3, 3), _ifield(ln, 6, 3)
        r0hb[i], phb1[i], phb2[i], phb3[i] = _floats(ln, 9, 4)
        inxn3hb[i1, i2, i3] = i
    eta = eta * 2.0                                          # :361
    loc = dict(locals())
    for name in ("rat rapt vnq Val Valboc mass Vale plp1 plp2 nlpopt povun2 povun3 povun4 povun5 povun6 "
                 "povun7 povun8 pval3 pval5 Valval Valangle chi eta gam gamij r0s r0p r0pp rvdW Dij alpij gamW "
                 "C_lg Re_lg rcore ecore acore pbo1 pbo2 pbo3 pbo4 pbo5 pbo6 bom pboc1 pboc2 pboc3 pboc4 pboc5 "
                 "Desig Depi Depipi pbe1 pbe2 povun1 ovc v13cor cBOp1 cBOp3 cBOp5 pbo2h pbo4h pbo6h switch "
                 "inxn2 inxn3 inxn3hb inxn4 theta00 pval1 pval2 pval4 pval6 pval7 pval8 pval9 pval10 ppen1 ppen2 "
                 "ppen3 ppen4 pcoa1 pcoa2 pcoa3 pcoa4 V1 V2 V3 ptor1 ptor2 ptor3 ptor4 pcot1 pcot2 r0hb phb1 "
                 "phb2 phb3").split():
        setattr(ff, name, loc[name])
    return ff
