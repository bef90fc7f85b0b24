"""Host-side mirror of the reference's hot-path interface, bound to librxmd_b200.so through ctypes.

The reference exposes four external Fortran subroutines with module-global state (no plugin API):

    QEq(atype, pos, q)                      src/qeq.F90:2
    FORCE(atype, pos, f, q)                 src/pot.F90:2
    COPYATOMS(imode, dr, atype, pos, v, f, q)   src/comm.F90:2   (MODE_MOVE is the host-visible call, src/main.F90:75)

`Engine` keeps the same names, argument meaning and array shapes (`pos(NBUFFER,3)` == numpy `[3, NBUFFER]`,
Fortran by-reference in/out semantics) and plays the role of `module atoms` for the state those routines share
(`NATOMS`, `PE(0:13)`, `astr`, `nstep_qeq`, `qsfp/qsfv`, `qs/qt`).  Errors follow the reference: the message it
would print, then a hard stop (here: RxmdError).  There is no CPU fallback: without the CUDA library or without
a B200 the constructor raises.
"""
from __future__ import annotations

import ctypes as C
import os
import numpy as np

from .binding import RxgConfig, RxgFF, RxgBox

MODE_COPY, MODE_MOVE, MODE_CPBK, MODE_QCOPY1, MODE_QCOPY2 = 1, 2, 3, 4, 5   # src/module.F90:38-39
HINT_ATOMS_ON_DEVICE, HINT_Q_ON_DEVICE, HINT_DEFER_POS, HINT_CHARGES_STAY = 1, 2, 4, 8   # include/rxmd_b200.h (rxg_hint)

_LIB = None


class RxmdError(RuntimeError):
    pass


def library_path():
    return os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "librxmd_b200.so")


def load_library():
    """dlopen the C-ABI library; raises if it was not built (no fallback path exists)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = library_path()
    if not os.path.exists(path):
        raise RxmdError(f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`")
    L = C.CDLL(path)
    dp, ip, vp = C.POINTER(C.c_double), C.POINTER(C.c_int), C.c_void_p
    L.rxg_create.argtypes = [C.POINTER(RxgConfig), C.POINTER(vp)]
    L.rxg_set_forcefield.argtypes = [vp, C.POINTER(RxgFF)]
    L.rxg_set_box.argtypes = [vp, C.POINTER(RxgBox)]
    L.rxg_comm_init.argtypes = [vp, C.c_int, C.c_int, vp]
    L.rxg_comm_unique_id.argtypes = [vp]
    L.rxg_comm_peer_halo.argtypes = [vp]
    L.rxg_destroy.argtypes = [vp]
    L.rxg_last_error.argtypes = [vp]
    L.rxg_last_error.restype = C.c_char_p
    L.rxg_qeq.argtypes = [vp, ip, dp, dp, dp, dp, dp, ip]
    L.rxg_pqeq.argtypes = [vp, ip, dp, dp, dp, dp, dp, dp, ip]
    L.rxg_spos_upload.argtypes = [vp, C.c_int, dp]
    L.rxg_spos_download.argtypes = [vp, C.c_int, dp]
    L.rxg_pqeq_skips.argtypes = [vp]
    L.rxg_pqeq_skips.restype = C.c_longlong
    L.rxg_force.argtypes = [vp, ip, dp, dp, dp, dp, dp, dp]
    L.rxg_move.argtypes = [vp, ip, dp, dp, dp, dp, dp, dp, dp, dp]
    L.rxg_fetch_bonds.argtypes = [vp, ip, dp]
    L.rxg_timers.argtypes = [vp, dp]
    L.rxg_state_upload.argtypes = [vp, C.c_int, dp, dp, dp, dp, dp, dp]
    L.rxg_md_run.argtypes = [vp, C.c_int, C.c_double, C.c_int, C.c_double, C.c_int]
    L.rxg_md_prime.argtypes = [vp]
    L.rxg_state_download.argtypes = [vp, ip, dp, dp, dp, dp, dp, dp, dp]
    L.rxg_md_observe.argtypes = [vp, dp, dp, dp, ip, dp]
    L.rxg_md_velocity_stats.argtypes = [vp, dp]
    L.rxg_md_velocity_affine.argtypes = [vp, dp, dp]
    L.rxg_debug_fetch.argtypes = [vp, C.c_char_p, vp, C.c_longlong, C.POINTER(C.c_longlong)]
    L.rxg_debug_spmv.argtypes = [vp, dp, dp, C.c_int, dp]
    L.rxg_hint.argtypes = [vp, C.c_int]
    L.rxg_it_timer.argtypes = [vp, dp]
    L.rxg_launch_count.argtypes = [vp]
    L.rxg_launch_count.restype = C.c_longlong
    _LIB = L
    return L


def _dp(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_double))


def broadcast_bytes(dist, payload: bytes, src=0) -> bytes:
    """Rank `src`'s byte string on every rank (used for the 128-byte ncclUniqueId; works with gloo and nccl)."""
    import torch
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor(list(payload), dtype=torch.uint8, device=dev)
    dist.broadcast(t, src=src)
    return bytes(t.cpu().tolist())


def rank_of_vid(vid, vprocs):
    """Sequential rank of a vector id, src/init.F90:97 (x fastest)."""
    return vid[0] + vid[1] * vprocs[0] + vid[2] * vprocs[0] * vprocs[1]


_F64 = {"atype", "q", "qst", "gst", "hsq", "val", "BO0", "BO1", "BO2", "BO3", "dln_BOp1", "dln_BOp2", "dln_BOp3", "dBOp",
        "A0", "A1", "A2", "A3", "delta", "deltap1", "deltap2", "nlp", "dDlp", "deltalp", "cdbnd", "ccbnd", "pos", "f", "v", "spos",
        "prow", "acc"}
_I64 = {"rowbeg", "rowend", "nnz", "uoff", "rowoff"}
_U8 = {"umask"}
_U16 = {"col16"}


class Engine:
    """One rank == one GPU.  Arrays passed to QEq/FORCE/COPYATOMS are the host's NBUFFER-capacity arrays."""

    def __init__(self, sysm, cfg: RxgConfig, rank=0):
        self.L = load_library()
        self.sys, self.cfg, self.rank = sysm, cfg, rank
        self.NBUFFER = cfg.nbuffer
        self.h = C.c_void_p()
        rc = self.L.rxg_create(C.byref(cfg), C.byref(self.h))
        self._chk(rc)
        self._chk(self.L.rxg_set_forcefield(self.h, C.byref(sysm.pff.struct)))
        self._chk(self.L.rxg_set_box(self.h, C.byref(sysm.boxes[rank].struct)))
        # module atoms state shared by the entry points
        self.NATOMS = 0
        self.PE = np.zeros(14)
        self.astr = np.zeros(6)
        self.nstep_qeq = 0
        nb = self.NBUFFER
        self.qsfp, self.qsfv, self.qs, self.qt = (np.zeros(nb) for _ in range(4))
        self.spos = np.zeros((3, nb)) if cfg.isPQEq else None   # spos(NBUFFER,3), src/init.F90:117-120

    # -- error convention of the reference: print 'ERROR: ...' and stop (src/main.F90:403-407, src/comm.F90:467-472)
    def _chk(self, rc):
        if rc != 0:
            msg = self.L.rxg_last_error(self.h).decode() if self.h else "rxg_create failed"
            raise RxmdError(f"[rc={rc}] {msg}")

    def close(self):
        if self.h:
            self.L.rxg_destroy(self.h)
            self.h = C.c_void_p()

    def comm_init_torch(self, dist):
        """NCCL communicator of the library: rank 0 creates the 128-byte ncclUniqueId, torch.distributed (any backend)
        broadcasts it -- the job MPI_Bcast does in the Fortran shim."""
        nranks, rank = dist.get_world_size(), dist.get_rank()
        buf = (C.c_ubyte * 128)()
        if rank == 0:
            self._chk(self.L.rxg_comm_unique_id(C.cast(buf, C.c_void_p)))
        raw = broadcast_bytes(dist, bytes(buf))
        buf2 = (C.c_ubyte * 128).from_buffer_copy(raw)
        self._chk(self.L.rxg_comm_init(self.h, rank, nranks, C.cast(buf2, C.c_void_p)))

    def hint(self, flags):
        """Promises about the arrays of the next entry-point call (rxg_hint): what the Fortran shim states inside the main loop."""
        self._chk(self.L.rxg_hint(self.h, int(flags)))

    def peer_halo(self):
        """True when the per-iteration ghost refreshes use peer-memory windows (NVLink stores) instead of NCCL send/recv."""
        return bool(self.L.rxg_comm_peer_halo(self.h) & 1)

    def peer_allreduce(self):
        return bool(self.L.rxg_comm_peer_halo(self.h) & 2)

    def host_arrays(self, rank_state):
        """Allocate the host's NBUFFER-capacity arrays (src/init.F90:110-114) from a rank's resident atoms."""
        nb = self.NBUFFER
        n = len(rank_state["atype"])
        atype, q = np.zeros(nb), np.zeros(nb)
        pos, v, f = np.zeros((3, nb)), np.zeros((3, nb)), np.zeros((3, nb))
        atype[:n] = rank_state["atype"]
        pos[:, :n] = rank_state["pos"]
        if rank_state.get("v") is not None:
            v[:, :n] = rank_state["v"]
        if rank_state.get("q") is not None:
            q[:n] = rank_state["q"]
        self.NATOMS = n
        return atype, pos, v, f, q

    # -- subroutine QEq(atype, pos, q), src/qeq.F90:2
    def QEq(self, atype, pos, q):
        n = C.c_int(self.NATOMS)
        it = C.c_int(0)
        self._chk(self.L.rxg_qeq(self.h, C.byref(n), _dp(atype), _dp(pos), _dp(q), _dp(self.qsfp), _dp(self.qsfv), C.byref(it)))
        self.nstep_qeq = it.value

    # -- subroutine PQEq(atype, pos, q), src/pqeq.F90:2 (module-global spos is relaxed at its end)
    def PQEq(self, atype, pos, q):
        if self.spos is None:
            raise RxmdError("ERROR: PQEq called without PQEq parameters (isPQEq = 0)")
        n = C.c_int(self.NATOMS)
        it = C.c_int(0)
        self._chk(self.L.rxg_pqeq(self.h, C.byref(n), _dp(atype), _dp(pos), _dp(q), _dp(self.spos), _dp(self.qsfp), _dp(self.qsfv),
                                  C.byref(it)))
        self.nstep_qeq = it.value

    def pqeq_skips(self):
        return int(self.L.rxg_pqeq_skips(self.h))

    # -- subroutine FORCE(atype, pos, f, q), src/pot.F90:2
    def FORCE(self, atype, pos, f, q):
        n = C.c_int(self.NATOMS)
        self._chk(self.L.rxg_force(self.h, C.byref(n), _dp(atype), _dp(pos), _dp(f), _dp(q), _dp(self.PE), _dp(self.astr)))

    # -- subroutine COPYATOMS(imode, dr, atype, pos, v, f, q), src/comm.F90:2: only MODE_MOVE is called by the host
    # (src/main.F90:75); the other modes run inside QEq/FORCE on the device.
    def COPYATOMS(self, imode, dr, atype, pos, v, f, q):
        if imode != MODE_MOVE:
            raise RxmdError(f"ERROR: imode doesn't match in COPYATOMS: {imode}")
        n = C.c_int(self.NATOMS)
        if self.spos is not None:   # spos migrates with the atom (src/comm.F90:153,165-167)
            self._chk(self.L.rxg_spos_upload(self.h, self.NATOMS, _dp(self.spos)))
        self._chk(self.L.rxg_move(self.h, C.byref(n), _dp(atype), _dp(pos), _dp(v), _dp(q), _dp(self.qs), _dp(self.qt),
                                  _dp(self.qsfp), _dp(self.qsfv)))
        self.NATOMS = n.value
        if self.spos is not None:
            self._chk(self.L.rxg_spos_download(self.h, self.NATOMS, _dp(self.spos)))

    # -- device-resident stepping
    def state_upload(self, atype, pos, v=None, q=None, qsfp=None, qsfv=None):
        self._chk(self.L.rxg_state_upload(self.h, self.NATOMS, _dp(atype), _dp(pos), _dp(v), _dp(q), _dp(qsfp), _dp(qsfv)))

    def md_prime(self):
        self._chk(self.L.rxg_md_prime(self.h))

    def md_run(self, nsteps, dt, qstep=1, Lex_w2=0.0, step0=0):
        self._chk(self.L.rxg_md_run(self.h, nsteps, dt, qstep, Lex_w2, step0))

    def state_download(self, atype, pos, v, f, q):
        n = C.c_int(0)
        self._chk(self.L.rxg_state_download(self.h, C.byref(n), _dp(atype), _dp(pos), _dp(v), _dp(f), _dp(q), _dp(self.qsfp),
                                            _dp(self.qsfv)))
        self.NATOMS = n.value

    def md_observe(self):
        ke, qs, it = C.c_double(), C.c_double(), C.c_int()
        astr = np.zeros(6)
        self._chk(self.L.rxg_md_observe(self.h, _dp(self.PE), C.byref(ke), C.byref(qs), C.byref(it), _dp(astr)))
        self.nstep_qeq = it.value
        return self.PE.copy(), ke.value, qs.value, it.value

    # -- thermostat hooks (the host's mdmode 4/5/7/8 logic drives them; src/main.F90:49-62,684-803)
    def velocity_stats(self):
        """Per atom type: count, sum 1/2 m v^2, sum m, sum m v (3) of this rank's residents -> array [nso, 6]."""
        nso = self.sys.pff.struct.nso
        out = np.zeros(6 * nso)
        self._chk(self.L.rxg_md_velocity_stats(self.h, _dp(out)))
        return out.reshape(nso, 6)

    def velocity_affine(self, scale, shift=(0.0, 0.0, 0.0)):
        """v(i) = scale[type(i)-1] * v(i) - shift on the resident velocities."""
        sc = np.ascontiguousarray(scale, dtype=np.float64)
        sh = np.ascontiguousarray(shift, dtype=np.float64)
        assert sc.size == self.sys.pff.struct.nso and sh.size == 3
        self._chk(self.L.rxg_md_velocity_affine(self.h, _dp(sc), _dp(sh)))

    def debug_spmv(self, x2, reps=1):
        """Row sums {H.x1, H.x2, ghost-column parts} of the production SpMV for x2[ntot, 2] (atom order); also its time."""
        ntot = int(self.fetch("copyptr")[6])
        n = int(self.fetch("copyptr")[0])
        x = np.ascontiguousarray(x2, dtype=np.float64).reshape(ntot, 2)
        out = np.zeros((n, 4))
        ms = C.c_double(0.0)
        self._chk(self.L.rxg_debug_spmv(self.h, _dp(x), _dp(out), int(reps), C.byref(ms)))
        return out, ms.value

    def natoms_resident(self):
        return int(self.fetch("copyptr")[0])

    def fetch_copyptr(self):
        return self.fetch("copyptr")

    def timers(self):
        t = np.zeros(30)
        self.L.rxg_timers(self.h, _dp(t))
        return t

    def it_timer(self):
        """The reference's it_timer(1:30) in seconds (index k-1 = slot k; slot 24 = QEq iterations)."""
        t = np.zeros(30)
        self._chk(self.L.rxg_it_timer(self.h, _dp(t)))
        return t

    def launches(self):
        return int(self.L.rxg_launch_count(self.h))

    def fetch(self, name):
        """Device -> host copy of a hot-path product (parity tests)."""
        cnt = C.c_longlong(0)
        self._chk(self.L.rxg_debug_fetch(self.h, name.encode(), None, 0, C.byref(cnt)))
        if name in _F64:
            out = np.empty(cnt.value, dtype=np.float64)
        elif name in _I64:
            out = np.empty(cnt.value, dtype=np.int64)
        elif name in _U8:
            out = np.empty(cnt.value, dtype=np.uint8)
        elif name in _U16:
            out = np.empty(cnt.value, dtype=np.uint16)
        else:
            out = np.empty(cnt.value, dtype=np.int32)
        if cnt.value:
            self._chk(self.L.rxg_debug_fetch(self.h, name.encode(), out.ctypes.data_as(C.c_void_p), out.nbytes, C.byref(cnt)))
        return out
