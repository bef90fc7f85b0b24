"""Structural minimiser (`mdmode 10`): the reference's second caller of the hot-path entry points.

Host-side restatement of `module CG`, reference src/cg.F90:26-390 (Polak-Ribiere conjugate gradient with a bracketing
step, Wolfe tests and a golden-section line search).  Like the Fortran original it is written against nothing but the
three entry points and the module state they share,

    QEq(atype, pos, q)   FORCE(atype, pos, f, q)   COPYATOMS(MODE_MOVE, 0, atype, pos, v, f, q)   NATOMS, PE(0:13)

so it runs unchanged on `Engine` (the CUDA library through the C ABI) -- which is the point: SURVEY 8f row 4 asks that the
minimiser work as a drop-in caller.  What it exercises beyond the MD loop: MODE_MOVE carrying an ARBITRARY 3-vector in the
`v` slot (search direction, gradient: `MigrateVec3D`, src/cg.F90:292-314), repeated evaluation on temporary arrays with
`NATOMS` saved and restored around the call (:367-381), and QEq started from a zero charge guess.

Multi-rank: `allreduce` is the reference's MPI_ALLREDUCE(SUM); the default is the single-rank identity.
"""
from __future__ import annotations

import numpy as np

from .engine import MODE_MOVE

CG_MaxMinLoop = 500          # src/cg.F90:5-8
CG_MaxBracketLoop = 20
CG_MaxLineMinLoop = 100
CG_WC1, CG_WC2 = 1e-4, 0.1   # :11-12
CG_GStol = 1e-6              # :15


class Minimizer:
    def __init__(self, backend, gnatoms, ftol=1e-4, allreduce=None, log=None, max_loops=CG_MaxMinLoop):
        self.b, self.gnatoms, self.ftol, self.max_loops = backend, gnatoms, ftol, max_loops
        self.allreduce = allreduce or (lambda x: x)
        self.log = log or (lambda *a: None)
        self.nb = backend.NBUFFER
        self.evaluations = 0
        self.history = []          # GPE(0) after every CG loop

    # -- helpers ---------------------------------------------------------------------------------------------------
    def _gpe(self):
        pe = np.array(self.b.PE, dtype=np.float64)
        pe[0] = pe[1:14].sum()                                  # PE(0)=sum(PE(1:13)), src/cg.F90:51
        return float(self.allreduce(pe)[0])

    def dot(self, a, c, n):                                     # DotProductVec3D, :318-335
        return float(self.allreduce(np.array([(a[:, :n] * c[:, :n]).sum()]))[0])

    def energy_with_step(self, atype, pos, p, stepl):           # EvaluateEnergyWithStep, :358-387
        b, n = self.b, self.b.NATOMS
        pos_t, atype_t = np.zeros((3, self.nb)), np.zeros(self.nb)
        pos_t[:, :n] = pos[:, :n] + stepl * p[:, :n]
        atype_t[:n] = atype[:n]
        vdummy, f_t, q_t = np.zeros((3, self.nb)), np.zeros((3, self.nb)), np.zeros(self.nb)
        b.COPYATOMS(MODE_MOVE, [0.0, 0.0, 0.0], atype_t, pos_t, vdummy, f_t, q_t)
        b.QEq(atype_t, pos_t, q_t)
        b.FORCE(atype_t, pos_t, f_t, q_t)
        b.NATOMS = n
        self.evaluations += 1
        return self._gpe()

    def migrate_vec(self, pos, vec, direction, stepl):          # MigrateVec3D, :292-314
        b, n = self.b, self.b.NATOMS
        pos_t, atype_d, q_d, f_d = np.zeros((3, self.nb)), np.zeros(self.nb), np.zeros(self.nb), np.zeros((3, self.nb))
        pos_t[:, :n] = pos[:, :n] + stepl * direction[:, :n]
        atype_d[:n] = self._atype[:n]     # the reference passes an uninitialised atypedummy; MODE_MOVE needs live types to keep atoms
        b.COPYATOMS(MODE_MOVE, [0.0, 0.0, 0.0], atype_d, pos_t, vec, f_d, q_d)
        new_n = b.NATOMS
        b.NATOMS = n
        return new_n

    def wolfe(self, atype, pos, p, stepl):                      # WolfeConditions, :144-208
        b, n = self.b, self.b.NATOMS
        q_t, fbefore = np.zeros(self.nb), np.zeros((3, self.nb))
        b.QEq(atype, pos, q_t)
        b.FORCE(atype, pos, fbefore, q_t)
        gpe_before = self._gpe()
        pos_t, atype_t, vdummy, fafter = np.zeros((3, self.nb)), np.zeros(self.nb), np.zeros((3, self.nb)), np.zeros((3, self.nb))
        pos_t[:, :n] = pos[:, :n] + stepl * p[:, :n]
        atype_t[:n] = atype[:n]
        b.COPYATOMS(MODE_MOVE, [0.0, 0.0, 0.0], atype_t, pos_t, vdummy, fafter, q_t)
        b.QEq(atype_t, pos_t, q_t)
        b.FORCE(atype_t, pos_t, fafter, q_t)
        gpe_after = self._gpe()
        self.evaluations += 2
        lower = gpe_after < gpe_before
        pdotdf = self.dot(p, fbefore, n)
        armijo = gpe_after <= gpe_before + pdotdf * CG_WC1 * stepl
        b.NATOMS = n                                             # :192-197: shift the search vector the same way
        pos_t[:, :n] = pos[:, :n] + stepl * p[:, :n]
        p_t = np.zeros((3, self.nb))
        p_t[:, :n] = p[:, :n]
        atype_t[:n] = atype[:n]
        b.COPYATOMS(MODE_MOVE, [0.0, 0.0, 0.0], atype_t, pos_t, p_t, vdummy, q_t)
        curvature = self.dot(p_t, fafter, b.NATOMS) >= CG_WC2 * pdotdf     # NATOMS is the post-move count here, like fafter's
        b.NATOMS = n
        return lower, armijo, curvature

    def bracket(self, atype, pos, p):                           # BracketSearchRange, :101-141
        stepl = 1e-2 / self.gnatoms
        for _ in range(CG_MaxBracketLoop):
            stepl *= 2
            pe = self.energy_with_step(atype, pos, p, stepl)
            lower, w1, w2 = self.wolfe(atype, pos, p, stepl)
            self.log("bracket", stepl, pe, lower, w1, w2)
            if (not w1) or (not w1):                             # sic, :128
                return stepl
        raise RuntimeError("bracket was not found")              # the reference saves the configuration and stops

    def golden_section(self, atype, pos, p, ax, dx):            # GoldenSectionSearch, :242-281
        ratio = 1.0 / 1.61803398875
        bx, cx = dx - (dx - ax) * ratio, ax + (dx - ax) * ratio
        pe_b, pe_c = self.energy_with_step(atype, pos, p, bx), self.energy_with_step(atype, pos, p, cx)
        for _ in range(CG_MaxLineMinLoop):
            if abs(ax - dx) <= CG_GStol / self.gnatoms:
                break
            if pe_b < pe_c:
                dx = cx
            else:
                ax = bx
            bx, cx = dx - (dx - ax) * ratio, ax + (dx - ax) * ratio
            pe_b, pe_c = self.energy_with_step(atype, pos, p, bx), self.energy_with_step(atype, pos, p, cx)
        return ax, dx

    def line_minimization(self, atype, pos, p, g, stepl):       # LineMinimization, :211-239
        b = self.b
        _, stepl = self.golden_section(atype, pos, p, 0.0, stepl)   # dx is in/out: the reference steps by the right boundary
        # Literal: `MigrateVec3D(pos,p,g,stepl)` binds vec=p, dir=g (:232 against the signature :292) -- the comment there says
        # "migrate gradient vector g", the call moves p along pos + stepl*g.  Kept as written.
        self.migrate_vec(pos, p, g, stepl)
        n = b.NATOMS
        pos[:, :n] = pos[:, :n] + stepl * p[:, :n]
        fdummy, qlocal = np.zeros((3, self.nb)), np.zeros(self.nb)   # `q` is a LOCAL array of LineMinimization (:221), not the caller's
        b.COPYATOMS(MODE_MOVE, [0.0, 0.0, 0.0], atype, pos, p, fdummy, qlocal)
        return stepl

    # -- ConjugateGradient, src/cg.F90:26-98 ----------------------------------------------------------------------------
    def run(self, atype, pos):
        b = self.b
        self._atype = atype
        q, gnew, gold, p = np.zeros(self.nb), np.zeros((3, self.nb)), np.zeros((3, self.nb)), np.zeros((3, self.nb))
        b.QEq(atype, pos, q)
        b.FORCE(atype, pos, gnew, q)
        n = b.NATOMS
        p[:, :n] = gnew[:, :n]
        gpe_new = self._gpe()
        self.history.append(gpe_new)
        stepl = self.bracket(atype, pos, p)
        for loop in range(self.max_loops):
            self.line_minimization(atype, pos, p, gnew, stepl)
            n = b.NATOMS
            gold[:, :n] = gnew[:, :n]
            b.QEq(atype, pos, q)
            b.FORCE(atype, pos, gnew, q)
            gpe_old, gpe_new = gpe_new, self._gpe()
            self.history.append(gpe_new)
            if abs(gpe_new - gpe_old) <= self.ftol * self.gnatoms:
                self.log("converged", loop)
                return True
            b1, b2, b3 = self.dot(gold, gold, n), self.dot(gnew, gnew, n), self.dot(gnew, gold, n)
            p[:, :n] = (b2 - b3) / b1 * p[:, :n] + gnew[:, :n]
            stepl = self.bracket(atype, pos, p)
        return False
