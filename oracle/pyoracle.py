"""ctypes front-end of the CPU oracle.  TEST INFRASTRUCTURE ONLY (tests/, smoke(), bench cpu_baseline).

Never imported by the product package `rxmd_b200`.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import numpy as np

from rxmd_b200.host.binding import RxgConfig, RxgFF, RxgBox

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    so = os.path.join(_HERE, "librxmd_oracle.so")
    src = os.path.join(_HERE, "rxmd_oracle.cpp")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
        L.orc_create.argtypes = [C.POINTER(RxgConfig), C.POINTER(RxgFF), C.POINTER(RxgBox), C.c_int, C.POINTER(C.c_void_p)]
        L.orc_destroy.argtypes = [C.c_void_p]
        L.orc_last_error.argtypes = [C.c_void_p]
        L.orc_last_error.restype = C.c_char_p
        L.orc_set_atoms.argtypes = [C.c_void_p, C.c_int, C.c_int, dp, dp, dp, dp, dp, dp]
        L.orc_set_corrected.argtypes = [C.c_void_p, C.c_int]
        L.orc_set_terms.argtypes = [C.c_void_p, C.c_int]
        L.orc_set_pqeq_stale.argtypes = [C.c_void_p, C.c_int]
        L.orc_natoms.argtypes = [C.c_void_p, C.c_int]
        L.orc_set_spos.argtypes = [C.c_void_p, C.c_int, dp]
        for f in (L.orc_qeq, L.orc_force, L.orc_move):
            f.argtypes = [C.c_void_p]
        L.orc_md_run.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_int, C.c_double, C.c_int]
        L.orc_get_f64.argtypes = [C.c_void_p, C.c_int, C.c_char_p, dp, C.c_longlong]
        L.orc_get_f64.restype = C.c_longlong
        L.orc_get_i32.argtypes = [C.c_void_p, C.c_int, C.c_char_p, ip, C.c_longlong]
        L.orc_get_i32.restype = C.c_longlong
        L.orc_observe.argtypes = [C.c_void_p, dp, dp, dp, ip]
        L.orc_timers.argtypes = [C.c_void_p, dp, dp, dp]
        L.orc_set_threads.argtypes = [C.c_int]
        _LIB = L
    return _LIB


def set_threads(n=0):
    """OpenMP team size of the oracle (n > 0 sets it); returns what the runtime will use."""
    return int(lib().orc_set_threads(int(n)))


def _dp(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_double))


class Oracle:
    """All ranks of one decomposition, simulated in-process."""

    def __init__(self, sysm, cfg: RxgConfig):
        self.L = lib()
        self.sys = sysm
        self.cfg = cfg
        self.nranks = len(sysm.boxes)
        boxes = (RxgBox * self.nranks)(*[b.struct for b in sysm.boxes])
        self._boxes = boxes
        self.h = C.c_void_p()
        rc = self.L.orc_create(C.byref(cfg), C.byref(sysm.pff.struct), boxes, self.nranks, C.byref(self.h))
        assert rc == 0
        for r, st in enumerate(sysm.ranks):
            self.set_atoms(r, st["atype"], st["pos"], st.get("v"), st.get("q"))

    def close(self):
        if self.h:
            self.L.orc_destroy(self.h)
            self.h = None

    def _chk(self, rc):
        if rc != 0:
            raise RuntimeError(f"oracle rc={rc}: {self.L.orc_last_error(self.h).decode()}")

    def set_atoms(self, rank, atype, pos, v=None, q=None, qsfp=None, qsfv=None):
        n = len(atype)
        a = [None if x is None else np.ascontiguousarray(x, dtype=np.float64) for x in (atype, pos, v, q, qsfp, qsfv)]
        self._chk(self.L.orc_set_atoms(self.h, rank, n, *[_dp(x) for x in a]))

    def set_spos(self, rank, spos):
        a = np.ascontiguousarray(spos, dtype=np.float64)
        self._chk(self.L.orc_set_spos(self.h, rank, _dp(a)))

    def set_corrected(self, on):
        self.L.orc_set_corrected(self.h, int(on))

    def set_pqeq_stale(self, on):
        self.L.orc_set_pqeq_stale(self.h, int(on))

    def set_terms(self, mask):
        self.L.orc_set_terms(self.h, int(mask))

    def qeq(self):
        self._chk(self.L.orc_qeq(self.h))

    def force(self):
        self._chk(self.L.orc_force(self.h))

    def move(self):
        self._chk(self.L.orc_move(self.h))

    def md_run(self, nsteps, dt, qstep=1, Lex_w2=0.0, step0=0):
        self._chk(self.L.orc_md_run(self.h, nsteps, dt, qstep, Lex_w2, step0))

    def natoms(self, rank=0):
        return self.L.orc_natoms(self.h, rank)

    def f64(self, name, rank=0):
        n = self.L.orc_get_f64(self.h, rank, name.encode(), None, 0)
        if n < 0:
            raise KeyError(name)
        out = np.empty(n)
        self.L.orc_get_f64(self.h, rank, name.encode(), _dp(out), n)
        return out

    def i32(self, name, rank=0):
        n = self.L.orc_get_i32(self.h, rank, name.encode(), None, 0)
        if n < 0:
            raise KeyError(name)
        out = np.empty(n, dtype=np.int32)
        self.L.orc_get_i32(self.h, rank, name.encode(), out.ctypes.data_as(C.POINTER(C.c_int)), n)
        return out

    def observe(self):
        pe = np.zeros(14)
        ke, qs, nq = C.c_double(), C.c_double(), C.c_int()
        self.L.orc_observe(self.h, _dp(pe), C.byref(ke), C.byref(qs), C.byref(nq))
        return pe, ke.value, qs.value, nq.value

    def timers(self):
        a, b, c = C.c_double(), C.c_double(), C.c_double()
        self.L.orc_timers(self.h, C.byref(a), C.byref(b), C.byref(c))
        return a.value, b.value, c.value
