"""ctypes mirror of include/rxmd_b200.h and the marshalling a Fortran host would do.

`pack_ff` / `pack_box` hand the library exactly what `gpu_shim.F90` passes from the Fortran
modules: 1-based arrays by their first element, column-major (see the header's conventions).
"""
from __future__ import annotations

import ctypes as C
import os
import numpy as np

from .ffield import ForceField
from . import setup as S

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int)


class RxgConfig(C.Structure):
    _fields_ = [("device", C.c_int), ("nbuffer", C.c_int), ("maxneighbs", C.c_int), ("maxneighbs10", C.c_int),
                ("nmincell", C.c_int), ("isQEq", C.c_int), ("NMAXQEq", C.c_int), ("isPQEq", C.c_int),
                ("isEfield", C.c_int), ("eFieldDir", C.c_int), ("QEq_tol", C.c_double), ("Lex_fqs", C.c_double),
                ("eFieldStrength", C.c_double)]


_FF_TYPE = "Val Valval Valangle Vale mass plp1 plp2 nlpopt povun2 povun3 povun4 povun5 povun6 povun7 povun8 pval3 pval5 chi eta".split()
_FF_BOND = ("cBOp1 cBOp3 cBOp5 pbo2h pbo4h pbo6h pbo2 pbo4 pbo6 swtch rc2 pboc1 pboc3 pboc4 pboc5 ovc v13cor "
            "Desig Depi Depipi pbe1 pbe2 povun1").split()
_FF_ANG = "theta00 pval1 pval2 pval4 pval6 pval7 pval8 pval9 pval10 ppen1 ppen2 ppen3 ppen4 pcoa1 pcoa2 pcoa3 pcoa4".split()
_FF_TOR = "ptor1 ptor2 ptor3 ptor4 V1 V2 V3 pcot1 pcot2".split()
_FF_HB = "phb1 phb2 phb3 r0hb".split()


class RxgFF(C.Structure):
    _fields_ = ([(n, C.c_int) for n in "nso nboty nvaty ntoty nhbty ntable".split()]
                + [(n, C.c_double) for n in "vpar1 vpar2 cutoff_vpar30 rctap rctap2 UDR UDRi".split()]
                + [(n, c_dp) for n in _FF_TYPE + _FF_BOND + _FF_ANG + _FF_TOR + _FF_HB]
                + [(n, c_ip) for n in "inxn2 inxn3 inxn3hb inxn4".split()]
                + [(n, c_dp) for n in "TBL_Evdw TBL_Eclmb TBL_Eclmb_QEq".split()]
                + [("ntype_pqeq", C.c_int), ("isPolarizable", c_ip), ("Zpqeq", c_dp), ("Kspqeq", c_dp),
                   ("inxnpqeq", c_ip), ("TBL_Eclmb_pcc", c_dp), ("TBL_Eclmb_psc", c_dp), ("TBL_Eclmb_pss", c_dp)])


class RxgBox(C.Structure):
    _fields_ = [("HH", C.c_double * 9), ("HHi", C.c_double * 9), ("lata", C.c_double), ("latb", C.c_double),
                ("latc", C.c_double), ("LBOX", C.c_double * 3), ("OBOX", C.c_double * 3), ("lcsize", C.c_double * 3),
                ("nblcsize", C.c_double * 3), ("cc", C.c_int * 3), ("nbcc", C.c_int * 3), ("nbnmesh", C.c_int),
                ("vprocs", C.c_int * 3), ("vID", C.c_int * 3), ("myparity", C.c_int * 3), ("target_node", C.c_int * 6),
                ("myid", C.c_int), ("nprocs", C.c_int), ("nbmesh", c_ip)]


def _dptr(a):
    return a.ctypes.data_as(c_dp)


def _iptr(a):
    return a.ctypes.data_as(c_ip)


class PackedFF:
    """Owns the contiguous buffers behind an RxgFF (keeps them alive)."""

    def __init__(self, ff: ForceField, rc2, tables, rctap, cutoff_vpar30, pqeq=None, chi=None, eta=None):
        T_vdw, T_clmb, T_qeq, UDR, UDRi = tables
        self.keep = {}
        st = RxgFF()
        st.nso, st.nboty, st.nvaty, st.ntoty, st.nhbty, st.ntable = ff.nso, ff.nboty, ff.nvaty, ff.ntoty, ff.nhbty, S.NTABLE
        st.vpar1, st.vpar2, st.cutoff_vpar30 = ff.vpar1, ff.vpar2, cutoff_vpar30
        st.rctap, st.rctap2, st.UDR, st.UDRi = rctap, rctap ** 2, UDR, UDRi

        def put(name, arr):
            a = np.ascontiguousarray(arr, dtype=np.float64)
            if a.size == 0:
                a = np.zeros(1)
            self.keep[name] = a
            setattr(st, name, _dptr(a))

        for n in _FF_TYPE + _FF_BOND + _FF_ANG + _FF_TOR + _FF_HB:
            if n == "swtch":
                put(n, np.asarray(ff.switch)[1:, 1:].ravel(order="F"))      # switch(1:3,1:nboty) column-major
            elif n == "rc2":
                put(n, np.asarray(rc2)[1:])
            elif n == "chi" and chi is not None:      # initialize_pqeq overwrites chi/eta (src/module.F90:517-523)
                put(n, np.asarray(chi)[1:])
            elif n == "eta" and eta is not None:
                put(n, np.asarray(eta)[1:])
            else:
                put(n, np.asarray(getattr(ff, n))[1:])
        for n, nd in (("inxn2", 2), ("inxn3", 3), ("inxn3hb", 3), ("inxn4", 4)):
            a = np.asarray(getattr(ff, n))[(slice(1, None),) * nd]
            a = np.ascontiguousarray(a.ravel(order="F"), dtype=np.int32)
            self.keep[n] = a
            setattr(st, n, _iptr(a))
        # TBL_Evdw(0:1,1:NTABLE,1:nboty) column-major from [c, i, inxn]
        put("TBL_Evdw", T_vdw[:, 1:, 1:].ravel(order="F"))
        put("TBL_Eclmb", T_clmb[:, 1:, 1:].ravel(order="F"))
        put("TBL_Eclmb_QEq", T_qeq[1:, 1:].ravel(order="F"))
        st.ntype_pqeq = 0
        if pqeq is not None:
            # module pqeq_vars as the Fortran host holds it: 1-based arrays by their first element, column-major
            n = pqeq.ntype
            st.ntype_pqeq = n
            pol = np.ascontiguousarray(np.asarray(pqeq.polarizable)[1:], dtype=np.int32)
            inx = np.ascontiguousarray(np.asarray(pqeq.inxnpqeq)[1:, 1:].ravel(order="F"), dtype=np.int32)
            self.keep["isPolarizable"], self.keep["inxnpqeq"] = pol, inx
            st.isPolarizable, st.inxnpqeq = _iptr(pol), _iptr(inx)
            put("Zpqeq", np.asarray(pqeq.Z)[1:])
            put("Kspqeq", np.asarray(pqeq.Ks)[1:])
            for nm, T in (("TBL_Eclmb_pcc", pqeq.T_pcc), ("TBL_Eclmb_psc", pqeq.T_psc), ("TBL_Eclmb_pss", pqeq.T_pss)):
                put(nm, T[1:, 1:, :].ravel(order="F"))   # (ntype_pqeq2, NTABLE, 0:1)
        self.struct = st


class PackedBox:
    def __init__(self, lattice, vprocs, myid, maxrc, rctap):
        la, lb, lc, al, be, ga = lattice
        H = S.get_box_params(la, lb, lc, al, be, ga)
        Hi = S.matinv(H)
        vID, parity, target = S.rank_topology(myid, vprocs)
        st = RxgBox()
        st.HH[:] = list(H.ravel(order="F"))
        st.HHi[:] = list(Hi.ravel(order="F"))
        st.lata, st.latb, st.latc = la, lb, lc
        lbox_real = [la / vprocs[0], lb / vprocs[1], lc / vprocs[2]]
        cc = [int(x / maxrc) for x in lbox_real]                              # src/init.F90:656
        LBOX = [1.0 / v for v in vprocs]
        st.LBOX[:] = LBOX
        st.lcsize[:] = [LBOX[a] / cc[a] for a in range(3)]
        st.OBOX[:] = [LBOX[a] * vID[a] for a in range(3)]
        st.cc[:] = cc
        nbcc, nblcsize, nbmesh = S.nonbonding_mesh(la, lb, lc, vprocs, rctap)
        st.nbcc[:] = [int(x) for x in nbcc]
        st.nblcsize[:] = list(nblcsize)
        self.nbmesh = np.ascontiguousarray(nbmesh, dtype=np.int32)           # [n,3] row-major == nbmesh(3,n) column-major
        st.nbnmesh = len(self.nbmesh)
        st.nbmesh = _iptr(self.nbmesh)
        st.vprocs[:] = list(vprocs)
        st.vID[:] = vID
        st.myparity[:] = parity
        st.target_node[:] = target
        st.myid = myid
        st.nprocs = int(np.prod(vprocs))
        self.struct = st
        self.H, self.Hi = H, Hi
        self.mdbox = float(np.linalg.det(H))


def repo_root():
    return os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
