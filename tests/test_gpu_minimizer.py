"""mdmode 10: the reference's structural minimiser (src/cg.F90) as a second caller of the same entry points.

The host algorithm (`rxmd_b200/host/minimize.py`, a restatement of `module CG`) is run twice -- once on the CUDA library through
the C ABI, once on the oracle through an adapter with the same three methods -- in the serial-order validation mode, where
the charges of the two are bit-identical, so every Wolfe / golden-section decision falls the same way and the two
minimisations can be compared call by call."""
import os

import numpy as np
import pytest

from rxmd_b200.host.system import build_system

pytestmark = pytest.mark.gpu

INP = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "inputs", "init.rdx")


class OracleBackend:
    """QEq / FORCE / COPYATOMS(MODE_MOVE) + NATOMS, PE over the CPU oracle (test infrastructure)."""

    def __init__(self, sysm, cfg):
        from oracle.pyoracle import Oracle
        self.o = Oracle(sysm, cfg)
        self.NBUFFER = cfg.nbuffer
        self.NATOMS = len(sysm.ranks[0]["atype"])
        self.PE = np.zeros(14)

    def _load(self, atype, pos, v=None, q=None):
        n = self.NATOMS
        self.o.set_atoms(0, atype[:n].copy(), pos[:, :n].copy(), None if v is None else v[:, :n].copy(), None if q is None else q[:n].copy())

    def QEq(self, atype, pos, q):
        n = self.NATOMS
        self._load(atype, pos, None, q)
        self.o.qeq()
        q[:n] = self.o.f64("q")[:n]
        pos[:, :n] = self.o.f64("pos").reshape(3, -1)[:, :n]

    def FORCE(self, atype, pos, f, q):
        n = self.NATOMS
        self._load(atype, pos, None, q)
        self.o.force()
        f[:, :n] = self.o.f64("f").reshape(3, -1)[:, :n]
        pos[:, :n] = self.o.f64("pos").reshape(3, -1)[:, :n]
        self.PE = self.o.f64("PE").copy()

    def COPYATOMS(self, imode, dr, atype, pos, v, f, q):
        assert imode == 2
        self._load(atype, pos, v, q)
        self.o.move()
        m = self.o.natoms()
        atype[:m] = self.o.f64("atype")[:m]
        pos[:, :m] = self.o.f64("pos").reshape(3, -1)[:, :m]
        v[:, :m] = self.o.f64("v").reshape(3, -1)[:, :m]
        q[:m] = self.o.f64("q")[:m]
        self.NATOMS = m


def test_minimiser_runs_on_the_entry_points_and_matches_the_oracle(built):
    from rxmd_b200.host.engine import Engine
    from rxmd_b200.host.minimize import Minimizer
    os.environ["RXG_STRICT_ORDER"] = "1"
    try:
        s = build_system(os.path.join(INP, "input.xyz"), os.path.join(INP, "ffield"), displace_sigma=0.05)
        cfg = s.config()
        n = s.natoms
        runs = {}
        for name in ("gpu", "oracle"):
            b = Engine(s, cfg) if name == "gpu" else OracleBackend(s, cfg)
            nb = cfg.nbuffer
            atype, pos = np.zeros(nb), np.zeros((3, nb))
            atype[:n] = s.ranks[0]["atype"]
            pos[:, :n] = s.ranks[0]["pos"]
            b.NATOMS = n
            calls = []
            m = Minimizer(b, gnatoms=n, ftol=1e-4, max_loops=2, log=lambda *a: calls.append(a))
            m.run(atype, pos)
            runs[name] = (m.history, calls, pos[:, :b.NATOMS].copy(), m.evaluations)
            if name == "gpu":
                b.close()
            else:
                b.o.close()
    finally:
        os.environ.pop("RXG_STRICT_ORDER", None)
    hg, cg, pg, eg = runs["gpu"]
    ho, co, po, eo = runs["oracle"]
    assert eg == eo and eg > 50                                   # the same number of energy evaluations (same decisions)
    assert len(hg) == len(ho) == 3
    assert hg[-1] < hg[0] - 1.0                                   # the displaced crystal relaxes (kcal/mol, 168 atoms)
    assert np.allclose(hg, ho, rtol=1e-10, atol=0.0)
    assert [c[0] for c in cg] == [c[0] for c in co]
    for a, b_ in zip(cg, co):
        if a[0] == "bracket":
            assert a[3:] == b_[3:] and abs(a[2] - b_[2]) <= 1e-10 * abs(b_[2])
    assert np.abs(pg - po).max() < 1e-8
