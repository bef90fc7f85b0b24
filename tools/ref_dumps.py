"""Write the hot path's results in the reference's own debug-dump formats, so that anyone with a Fortran toolchain can diff
them against the real binary built with -DRFDUMP / -DQEQDUMP (src/pot.F90:76-88, src/qeq.F90:65-112):

  rfdump<rank>.txt   per resident, three blocks 'pos', 'frc', 'chg':  format (i6,1x,a3,i6,7f20.12): gid, tag, type, values
  qeqdump<rank>.txt  per stored pair of the QEq matrix:               format (4i6,4es25.15): -1, gid(i), type(i), gid(j), hessian

    python tools/ref_dumps.py --config rdx --mc 1 1 1 [--out DIR] [--cpu]

--cpu writes the same files from the CPU oracle instead (test infrastructure) -- the pair of outputs is what the parity tests
compare in memory.  Rank numbering and file names follow rankToString(myid).
"""
import argparse
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np


def fortran_f(x, w=20, d=12):
    return f"{x:{w}.{d}f}"


def fortran_es(x, w=25, d=15):
    # Fortran ES25.15: d.ddddddddddddddde+XX (two-digit exponent)
    s = f"{x:.{d}E}"
    m, e = s.split("E")
    return f"{m}E{int(e):+03d}".rjust(w)


def write_rfdump(path, gid, ity, pos, f, q):
    with open(path, "w") as fh:
        for tag, arr in (("pos", pos), ("frc", f)):
            for i in range(len(gid)):
                fh.write(f"{gid[i]:6d} {tag}{ity[i]:6d}" + "".join(fortran_f(arr[c, i]) for c in range(3)) + "\n")
        for i in range(len(gid)):
            fh.write(f"{gid[i]:6d} chg{ity[i]:6d}" + fortran_f(q[i]) + "\n")


def write_qeqdump(path, gid_all, ity, rowbeg, rowend, col, val):
    with open(path, "w") as fh:
        for i in range(len(rowbeg)):
            for k in range(rowbeg[i], rowend[i]):
                fh.write(f"{-1:6d}{gid_all[i]:6d}{ity[i]:6d}{gid_all[col[k]]:6d}" + fortran_es(val[k]) + "\n")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="rdx")
    ap.add_argument("--mc", type=int, nargs=3, default=[1, 1, 1])
    ap.add_argument("--sigma", type=float, default=0.0)
    ap.add_argument("--out", default=".")
    ap.add_argument("--cpu", action="store_true", help="dump the CPU oracle's results instead of the CUDA library's")
    a = ap.parse_args()
    from rxmd_b200.host.configs import build_config
    s, tot, vp, cfgkw, label = build_config(a.config, mc=a.mc, sigma=a.sigma)
    cfg = s.config(**cfgkw)
    os.makedirs(a.out, exist_ok=True)
    st = s.ranks[0]
    n = len(st["atype"])
    if a.cpu:
        from oracle.pyoracle import Oracle
        o = Oracle(s, cfg)
        o.qeq()
        atype_all = o.f64("atype")
        W = cfg.maxneighbs10
        cnt, lst, hes = o.i32("nbpcnt"), o.i32("nbplist").reshape(n, W), o.f64("hessian").reshape(n, W)
        rb = np.concatenate([[0], np.cumsum(cnt)])[:-1]
        col = np.concatenate([lst[i, :cnt[i]] for i in range(n)])
        val = np.concatenate([hes[i, :cnt[i]] for i in range(n)])
        re_ = rb + cnt
        o.force()
        pos, f, q = o.f64("pos").reshape(3, -1)[:, :n], o.f64("f").reshape(3, -1)[:, :n], o.f64("q")[:n]
        o.close()
    else:
        from rxmd_b200.host.engine import Engine
        e = Engine(s, cfg)
        atype, pos, v, f, q = e.host_arrays(st)
        (e.PQEq if cfg.isPQEq else e.QEq)(atype, pos, q)
        atype_all = e.fetch("atype")
        rb, re_, col, val = e.fetch("rowbeg"), e.fetch("rowend"), e.fetch("col"), e.fetch("val")
        e.FORCE(atype, pos, f, q)
        pos, f, q = pos[:, :n], f[:, :n], q[:n]
        e.close()
    ity_all = np.rint(atype_all).astype(int)
    gid_all = np.rint((atype_all - ity_all) * 1e13).astype(int)      # l2g(atype), src/main.F90:582-593
    write_qeqdump(os.path.join(a.out, "qeqdump0.txt"), gid_all, ity_all, rb, re_, col, val)
    write_rfdump(os.path.join(a.out, "rfdump0.txt"), gid_all[:n], ity_all[:n], pos, f, q)
    print(f"wrote {a.out}/rfdump0.txt ({3 * n} lines) and {a.out}/qeqdump0.txt ({int((re_ - rb).sum())} lines) for {label} x{tot}")


if __name__ == "__main__":
    main()
