"""Stage-by-stage GPU-vs-oracle comparison (development diagnostic; run on the GPU box).

    python tools/gpu_diag.py [mcx mcy mcz] [--sigma S]
"""
import sys
import os
import time
import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rxmd_b200.host.system import build_system  # noqa: E402
from rxmd_b200.host.engine import Engine  # noqa: E402
from oracle.pyoracle import Oracle  # noqa: E402

G = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests/golden/inputs/init.rdx/")


def rel(a, b):
    d = np.abs(a - b).max() if a.size else 0.0
    s = max(np.abs(b).max() if b.size else 0.0, 1e-300)
    return d, d / s


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    mc = tuple(int(x) for x in args[:3]) if len(args) >= 3 else (1, 1, 1)
    sigma = float(sys.argv[sys.argv.index("--sigma") + 1]) if "--sigma" in sys.argv else 0.0
    s = build_system(G + "input.xyz", G + "ffield", mc=mc, displace_sigma=sigma)
    cfg = s.config()
    print("mc", mc, "natoms", s.natoms, "nbuffer", cfg.nbuffer, flush=True)
    o = Oracle(s, cfg)
    e = Engine(s, cfg)
    st = s.ranks[0]
    atype, pos, v, f, q = e.host_arrays(st)
    if sigma > 0:   # exercise MODE_MOVE (displaced atoms are already wrapped by build_system)
        o.move()
        e.COPYATOMS(2, [0, 0, 0], atype, pos, v, f, q)
        n = e.NATOMS
        print("after move: natoms gpu", n, "oracle", o.natoms())
        print("  pos diff", rel(pos[:, :n], o.f64("pos").reshape(3, -1)[:, :n]), "atype equal", np.array_equal(atype[:n], o.f64("atype")[:n]))
    n = e.NATOMS
    # ---------------- QEq
    t0 = time.time(); o.qeq(); t_o = time.time() - t0
    t0 = time.time(); e.QEq(atype, pos, q); t_g = time.time() - t0
    print(f"QEq: oracle {t_o:.3f}s gpu {t_g:.3f}s  nstep_qeq oracle {o.observe()[3]} gpu {e.nstep_qeq}")
    cp_o, cp_g = o.i32("copyptr"), e.fetch("copyptr")
    print("  copyptr oracle", cp_o, "gpu", cp_g)
    n6 = cp_o[6]
    if np.array_equal(cp_o, cp_g):
        print("  ghost pos diff", rel(e.fetch("pos").reshape(3, -1), o.f64("pos").reshape(3, -1)))
        print("  ghost atype equal", np.array_equal(e.fetch("atype"), o.f64("atype")))
    cnt_o = o.i32("nbpcnt")
    rb, re_ = e.fetch("rowbeg"), e.fetch("rowend")
    cnt_g = re_ - rb
    print("  row counts equal", np.array_equal(cnt_o, cnt_g), "nnz (true)", int(cnt_g.sum()), cnt_o.sum(), "padded", e.fetch("nnz")[0])
    W = cfg.maxneighbs10
    lst_o = o.i32("nbplist").reshape(n, W)
    hes_o = o.f64("hessian").reshape(n, W)
    col, val = e.fetch("col"), e.fetch("val")
    same_order, same_set, hmax = True, True, 0.0
    for i in range(n):
        a = lst_o[i, :cnt_o[i]]
        b = col[rb[i]:re_[i]]
        if len(a) != len(b):
            same_order = same_set = False
            continue
        if not np.array_equal(a, b):
            same_order = False
            if not np.array_equal(np.sort(a), np.sort(b)):
                same_set = False
            else:
                hmax = max(hmax, np.abs(hes_o[i, :cnt_o[i]][np.argsort(a)] - val[rb[i]:re_[i]][np.argsort(b)]).max())
        else:
            hmax = max(hmax, np.abs(hes_o[i, :cnt_o[i]] - val[rb[i]:re_[i]]).max())
    print("  pair rows: same order", same_order, "same sets", same_set, "hessian max abs diff", hmax)
    print("  q diff (abs, rel)", rel(q[:n], o.f64("q")[:n]), " sum q", q[:n].sum())
    # ---------------- FORCE (identical charges on both sides: the oracle's)
    q[:n] = o.f64("q")[:n]
    t0 = time.time(); o.force(); t_o = time.time() - t0
    t0 = time.time(); e.FORCE(atype, pos, f, q); t_g = time.time() - t0
    print(f"FORCE: oracle {t_o:.3f}s gpu {t_g:.3f}s")
    cp_o, cp_g = o.i32("copyptr"), e.fetch("copyptr")
    print("  copyptr oracle", cp_o, "gpu", cp_g)
    n6 = cp_o[6]
    M = cfg.maxneighbs
    nc_o, nc_g = o.i32("nbrcnt"), e.fetch("nbrcnt")
    print("  nbrcnt equal", np.array_equal(nc_o, nc_g), "mean", nc_o.mean(), "max", nc_o.max())
    nl_o, nl_g = o.i32("nbrlist").reshape(n6, M), e.fetch("nbrlist").reshape(n6, M)
    mask = np.arange(M)[None, :] < nc_o[:, None]
    print("  nbrlist identical (ordered)", np.array_equal(nl_o[mask], nl_g[mask]))
    ni_o, ni_g = o.i32("nbrindx").reshape(n6, M), e.fetch("nbrindx").reshape(n6, M)
    print("  nbrindx identical", np.array_equal(ni_o[mask], ni_g[mask]))
    for name in ("BO0", "BO1", "BO2", "BO3", "dBOp", "dln_BOp1", "dln_BOp2", "dln_BOp3", "A0", "A1", "A2", "A3"):
        a, b = e.fetch(name).reshape(n6, M)[mask], o.f64(name).reshape(n6, M)[mask]
        print(f"  {name:9s} max abs/rel diff", rel(a, b))
    for name in ("deltap1", "deltap2", "delta", "nlp", "dDlp", "deltalp"):
        print(f"  {name:9s} max abs/rel diff", rel(e.fetch(name)[:n6], o.f64(name)[:n6]))
    pe_o = o.f64("PE")
    print("  PE oracle", np.array2string(pe_o[1:], precision=10))
    print("  PE gpu   ", np.array2string(e.PE[1:], precision=10))
    print("  PE rel diff", np.abs(e.PE[1:] - pe_o[1:]) / np.maximum(np.abs(pe_o[1:]), 1e-300))
    f_o = o.f64("f").reshape(3, -1)[:, :n]
    d, r = rel(f[:, :n], f_o)
    print("  f max abs diff", d, "relative to max|f|", r, " sum f gpu", f[:, :n].sum(axis=1))
    print("  astr oracle", o.f64("astr"), "gpu", e.astr)
    print("  launches", e.launches(), "timers(ms) qeq/force/move", e.timers()[:3])
    # ---------------- a few MD steps, device resident
    UTIME = 1e3 / 20.455
    dt = 0.25 / UTIME
    Lw2 = 2.0 * 2.0 / dt / dt
    e.state_upload(atype, pos, v, q)
    e.md_prime()
    o.qeq(); o.force()
    t0 = time.time(); e.md_run(10, dt, 1, Lw2, 0); t_g = time.time() - t0
    t0 = time.time(); o.md_run(10, dt, 1, Lw2, 0); t_o = time.time() - t0
    pe_g, ke_g, qs_g, it_g = e.md_observe()
    pe_o, ke_o, qs_o, it_o = o.observe()
    print(f"MD 10 steps: oracle {t_o:.3f}s gpu {t_g:.3f}s")
    print("  PE/atom gpu", pe_g[0] / n, "oracle", pe_o[0] / n, " KE/atom gpu", ke_g / n, "oracle", ke_o / n, " nstep_qeq", it_g, it_o)
    e.state_download(atype, pos, v, f, q)
    nn = e.NATOMS
    print("  natoms", nn, o.natoms(), " pos diff", rel(pos[:, :nn], o.f64("pos").reshape(3, -1)[:, :nn]),
          " v diff", rel(v[:, :nn], o.f64("v").reshape(3, -1)[:, :nn]))


if __name__ == "__main__":
    main()
