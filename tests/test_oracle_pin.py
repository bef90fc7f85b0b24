"""The oracle against the reference's only known-answer data (README.md:127-167) and against physics invariants.

README.md:157 (168-atom RDX, nitramine ffield, step 0, per-atom kcal/mol):
  TE=PE -9.82464E+01, Ebond -1.369E+02, Elp+Eover+Eunder 1.287E+00, Eval+Epen+Ecoa -1.362E+00,
  Etors+Econj 5.208E-01, Ehbond -1.398E-03, Evdw+Eclmb+Echarge 3.821E+01, sum q 0.00
Header of the same run: cells 4 3 3, maxrc 3.160, atoms per type 24/48/48/48 (README.md:139-152).
"""
import numpy as np
import pytest

from rxmd_b200.host.system import build_system
from rxmd_b200.host import setup as S


@pytest.fixture(scope="module")
def rdx(built, rdx_paths):
    from oracle.pyoracle import Oracle
    s = build_system(*rdx_paths)
    o = Oracle(s, s.config(nbuffer=30000))
    o.qeq()
    o.force()
    yield s, o
    o.close()


def test_static_pins(rdx):
    s, o = rdx
    b = s.boxes[0].struct
    assert s.natoms == 168
    assert list(b.cc) == [4, 3, 3] and list(b.nbcc) == [4, 3, 3]
    assert abs(s.maxrc - 3.16) < 1e-9
    ty = np.rint(s.ranks[0]["atype"]).astype(int)
    assert [int((ty == t).sum()) for t in (1, 2, 3, 4)] == [24, 48, 48, 48]
    assert b.nbnmesh == 305
    assert o.i32("copyptr")[6] == 168 * 27                    # 27 periodic images in the FORCE halo


def test_readme_step0_energies(rdx):
    s, o = rdx
    pe = o.observe()[0] / s.natoms

    def digits(x, ref, nd):       # agreement to the README's printed significant digits
        return abs(x - ref) <= 0.5 * 10 ** (np.floor(np.log10(abs(ref))) - nd + 1) * 1.0001

    assert digits(pe[0], -9.82464e1, 6)
    assert digits(pe[1], -1.369e2, 4)
    assert digits(pe[2:5].sum(), 1.287, 4)
    assert digits(pe[5:8].sum(), -1.362, 4)
    assert digits(pe[8:10].sum(), 5.208e-1, 4)
    assert digits(pe[10], -1.398e-3, 4)
    assert digits(pe[11:14].sum(), 3.821e1, 4)


def test_invariants(rdx):
    s, o = rdx
    n = s.natoms
    f = o.f64("f").reshape(3, -1)[:, :n]
    q = o.f64("q")[:n]
    assert np.abs(f.sum(axis=1)).max() < 1e-9 * np.abs(f).max() * n     # Newton's third law after the copy-back
    assert abs(q.sum()) < 1e-10                                          # README: sum q = 0.00
    cnt, lst, idx = o.i32("nbrcnt"), o.i32("nbrlist").reshape(-1, 30), o.i32("nbrindx").reshape(-1, 30)
    for i in range(0, len(cnt), 97):                                     # nbrlist <-> nbrindx consistency, src/main.F90:394
        for s1 in range(cnt[i]):
            j = lst[i, s1]
            assert lst[j, idx[i, s1]] == i


def test_tables_are_self_consistent(rdx_paths):
    """TBL(1) must be (dE/dr)/r of TBL(0) (src/init.F90:445-494): central differences in r^2."""
    s = build_system(*rdx_paths)
    T_vdw, T_clmb, T_qeq, UDR, UDRi = S.potential_table(s.ff, 10.0, S.taper(10.0))
    for inxn in (1, 2, 5):
        for T in (T_vdw, T_clmb):
            E, F = T[0, 1:, inxn], T[1, 1:, inxn]
            num = (E[2:] - E[:-2]) / (2 * UDR) * 2
            k = np.array([200, 1000, 3000, 4500])
            assert np.allclose(num[k - 1], F[k], rtol=2e-4)


def test_lg_tables_are_self_consistent():
    """The tables of the HEADLINE configuration (conf/init.rdx.lg, `isLG`): the low-gradient dispersion and inner-core terms the
    LG force field adds to TBL_Evdw (src/init.F90:496-512) have no reference numbers to be pinned against, so they are held to
    what the reference's own formulas must satisfy: TBL(1) = (dE/dr)/r of TBL(0) -- which checks dElg and dE_core against Elg and
    E_core -- and the LG part must vanish from the Coulomb / QEq tables and be the exact difference to the tables of the same
    file parsed without the LG columns' effect (C_lg = ecore = 0)."""
    import os
    from conftest import INPUTS
    d = os.path.join(INPUTS, "init.rdx.lg")
    s = build_system(os.path.join(d, "input.xyz"), os.path.join(d, "ffield"), isLG=True)
    assert s.ff.isLG
    T_vdw, T_clmb, T_qeq, UDR, UDRi = S.potential_table(s.ff, 10.0, S.taper(10.0))
    k = np.array([200, 1000, 3000, 4500])
    checked = 0
    for inxn in range(1, T_vdw.shape[2]):
        E, F = T_vdw[0, 1:, inxn], T_vdw[1, 1:, inxn]
        if not np.any(E):
            continue
        num = (E[2:] - E[:-2]) / (2 * UDR) * 2
        assert np.allclose(num[k - 1], F[k], rtol=2e-4, atol=1e-9), inxn
        checked += 1
    assert checked >= 6                                   # C, H, O, N pairs of RDX
    # switching the LG parameters off changes TBL_Evdw only, by exactly Tap*(Elg + E_core)
    import copy
    ff0 = copy.deepcopy(s.ff)
    ff0.C_lg = np.zeros_like(ff0.C_lg); ff0.ecore = np.zeros_like(ff0.ecore)
    V0, C0, Q0, _, _ = S.potential_table(ff0, 10.0, S.taper(10.0))
    assert np.array_equal(C0, T_clmb) and np.array_equal(Q0, T_qeq)
    assert np.abs(V0 - T_vdw).max() > 1e-3                # the LG terms are not negligible for this field


@pytest.mark.parametrize("field", ["nitramine", "lg"])
def test_bonded_forces_match_finite_differences(built, rdx_paths, field):
    """Derivative chain of every bonded term (Ebond, Elnpr, Ehb, E3b, E4b) in the oracle's `corrected` mode, where no
    ccbnd contribution is discarded; the literal mode differs from it by exactly those terms.  Run on the
    README's nitramine cell and on the LG cell of the headline configuration (conf/init.rdx.lg)."""
    from oracle.pyoracle import Oracle
    if field == "lg":
        import os
        from conftest import INPUTS
        d = os.path.join(INPUTS, "init.rdx.lg")
        s = build_system(os.path.join(d, "input.xyz"), os.path.join(d, "ffield"), isLG=True, displace_sigma=0.03)
    else:
        s = build_system(*rdx_paths, displace_sigma=0.03)
    o = Oracle(s, s.config(nbuffer=30000))
    o.move()
    n = o.natoms()
    at = o.f64("atype")[:n].copy()
    pos0 = o.f64("pos").reshape(3, -1)[:, :n].copy()
    o.qeq()
    q = o.f64("q")[:n].copy()
    o.set_corrected(1)
    h = 1e-5
    for mask in (2, 4, 8, 16, 32):
        o.set_terms(mask)
        o.set_atoms(0, at, pos0, None, q); o.force()
        f = o.f64("f").reshape(3, -1)[:, :n].copy()
        scale = max(np.abs(f).max(), 1e-3)
        for c, i in [(0, 0), (1, 5), (2, 30), (0, 60), (1, 100), (2, 150)]:
            p = pos0.copy(); p[c, i] += h
            o.set_atoms(0, at, p, None, q); o.force(); ep = o.observe()[0][0]
            p = pos0.copy(); p[c, i] -= h
            o.set_atoms(0, at, p, None, q); o.force(); em = o.observe()[0][0]
            fd = -(ep - em) / (2 * h)
            assert abs(fd - f[c, i]) < 2e-6 * scale + 1e-7, (mask, c, i, fd, f[c, i])
    o.close()


def test_decomposition_invariance(built, rdx_paths):
    """examples/1-reaxff vs examples/2-reaxff-dc in the reference: the same crystal on 1 and on 2 ranks gives the same
    per-atom energies to the QEq tolerance.  Forces: the reference's ForceBondedTerms discards some ccbnd
    contributions depending on LOCAL index order, so literal forces depend on the decomposition at
    the 10% level; in the oracle's `corrected` mode (nothing discarded) they agree to round-off, which is what checks the
    multi-rank COPYATOMS / copy-back machinery."""
    from oracle.pyoracle import Oracle
    s1 = build_system(*rdx_paths, mc=(2, 1, 1), vprocs=(1, 1, 1))
    s2 = build_system(*rdx_paths, mc=(2, 1, 1), vprocs=(2, 1, 1))
    o1, o2 = Oracle(s1, s1.config(nbuffer=30000)), Oracle(s2, s2.config(nbuffer=30000))
    o1.qeq(); o1.force()
    o2.qeq(); o2.force()
    pe1, pe2 = o1.observe()[0], o2.observe()[0]
    assert np.allclose(pe1[1:11], pe2[1:11], rtol=1e-10, atol=1e-9)          # bonded terms do not depend on q
    assert abs(pe1[0] - pe2[0]) / abs(pe1[0]) < 1e-6                          # total: QEq-tolerance level
    # forces with identical charges: map by global id
    g1 = np.rint((s1.ranks[0]["atype"] - np.rint(s1.ranks[0]["atype"])) * 1e13).astype(int)
    q_by_gid = dict(zip(g1, o1.f64("q")[:s1.natoms]))
    f_by_gid = {}
    for r in range(2):
        at = s2.ranks[r]["atype"]
        gid = np.rint((at - np.rint(at)) * 1e13).astype(int)
        o2.set_atoms(r, at, s2.ranks[r]["pos"], None, np.array([q_by_gid[g] for g in gid]))
    at1 = s1.ranks[0]["atype"]
    o1.set_atoms(0, at1, s1.ranks[0]["pos"], None, o1.f64("q")[:s1.natoms].copy())
    lit = []
    for corrected in (0, 1):
        o1.set_corrected(corrected); o2.set_corrected(corrected)
        o1.force(); o2.force()
        f1 = o1.f64("f").reshape(3, -1)[:, :s1.natoms]
        for r in range(2):
            n = len(s2.ranks[r]["atype"])
            at = s2.ranks[r]["atype"]
            gid = np.rint((at - np.rint(at)) * 1e13).astype(int)
            f2 = o2.f64("f", r).reshape(3, -1)[:, :n]
            for k, g in enumerate(gid):
                f_by_gid[g] = f2[:, k].copy()
        lit.append(max(np.abs(f1[:, k] - f_by_gid[g]).max() for k, g in enumerate(g1)) / np.abs(f1).max())
    assert lit[1] < 1e-9, lit
    assert lit[0] > 1e-6, lit            # Q1 is real: the literal algorithm is decomposition dependent
    o1.close(); o2.close()
