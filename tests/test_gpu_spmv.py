"""Kernel-level parity of the CG's sparse product (the kernel bench.py's roofline is quoted on), through the C-ABI.

`rxg_debug_spmv` runs exactly the launch the production CG issues -- k_spmv_rows, whose three launch shapes and unstaged
path are forced in turn through the RXG_SPMV* switches; RXG_SPMV=items selects the experimental cell-blocked kernel
k_spmv_items (full-size ring, small ring, unstaged walk) -- on a given vector pair and returns the four raw row sums per
resident row.  They are compared with an extended-precision NumPy product of the SAME matrix, fetched from the device and
already proven bit-identical to the oracle's hessian / nbplist (reference src/qeq.F90:222-240) in test_gpu_parity; the
bar is 1e-13 of sum_j |H_ij x_j|, i.e. summation-order round-off only.  The union stream that k_spmv_items walks is checked
structurally against the rows it was built from.
"""
import os

import numpy as np
import pytest

from rxmd_b200.host.system import build_system

pytestmark = pytest.mark.gpu

INP = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "inputs")
SPMV_ENV = ("RXG_SPMV", "RXG_SPMV_SHAPE", "RXG_SPMV_STAGE", "RXG_SPMV_RING", "RXG_FUSE_API", "RXG_WIN_G", "RXG_WIN_WARPS", "RXG_WIN_WCAP", "RXG_WIN_U", "RXG_WIN_RALIGN", "RXG_WIN_LPR",
            "RXG_WIN_SMEM")

VARIANTS = {
    # k_spmv_win (production): window of x in shared memory, 16-bit window-relative columns
    "win_auto": {},
    "win_g1_w4": {"RXG_WIN_G": "1", "RXG_WIN_WARPS": "4"},
    "win_g5_w4_u4": {"RXG_WIN_G": "5", "RXG_WIN_WARPS": "4", "RXG_WIN_U": "4"},
    "win_g3_u4": {"RXG_WIN_G": "3", "RXG_WIN_U": "4"},                     # short rows: a full warp per row, 4 entries per lane
    "win_g3_u4_lpr16": {"RXG_WIN_G": "3", "RXG_WIN_U": "4", "RXG_WIN_LPR": "16"},   # short rows: two rows per warp, 16 lanes each
    "win_g7": {"RXG_WIN_G": "7"},
    "win_ralign16": {"RXG_WIN_RALIGN": "16"},                   # rows on 128-byte boundaries of the value stream
    "win_fallback": {"RXG_WIN_G": "2", "RXG_WIN_WCAP": "512"},   # no window fits 512 entries: every CTA gathers from global memory
    "win_shared_list": {"RXG_FUSE_API": "1"},
    "items": {"RXG_SPMV": "items"},
    "items_3stage": {"RXG_SPMV": "items", "RXG_SPMV_RING": "3"},
    "items_unstaged": {"RXG_SPMV": "items", "RXG_SPMV_STAGE": "0"},
    "items_shared_list": {"RXG_SPMV": "items", "RXG_FUSE_API": "1"},      # the list rxg_md_run / bench.py use: FORCE predicate, zero hessian entries
    "rows_auto": {"RXG_SPMV": "rows"},
    "rows_8x8": {"RXG_SPMV": "rows", "RXG_SPMV_SHAPE": "8x8"},
    "rows_4x16": {"RXG_SPMV": "rows", "RXG_SPMV_SHAPE": "4x16"},
    "rows_2x32": {"RXG_SPMV": "rows", "RXG_SPMV_SHAPE": "2x32"},
    "rows_unstaged": {"RXG_SPMV": "rows", "RXG_SPMV_STAGE": "0"},
}


def systems():
    from test_gpu_parity import systems as base
    s = dict(base())
    pe = os.path.join(INP, "init.pe.pqeq")
    # polyethylene with the 12.5 A PQEq cut-off: ~1060 entries per row, the long-row case (k_spmv_rows<2,32,1216>)
    s["pe_pqeq_4x6x11"] = dict(xyz=os.path.join(pe, "input.xyz"), ff=os.path.join(pe, "ffield"), mc=(4, 6, 11), displace_sigma=0.02,
                               pqeq_path=os.path.join(pe, "pqeq1.par"))
    return s


@pytest.fixture(autouse=True)
def _clean_env():
    for k in SPMV_ENV:
        os.environ.pop(k, None)
    yield
    for k in SPMV_ENV:
        os.environ.pop(k, None)


def _engine(name, env):
    from rxmd_b200.host.engine import Engine
    os.environ.update(env)
    kw = dict(systems()[name])
    s = build_system(kw.pop("xyz"), kw.pop("ff"), **kw)
    cfg = s.config(NMAXQEq=2)                     # two CG iterations are enough to leave the matrix on the device
    e = Engine(s, cfg)
    atype, pos, v, f, q = e.host_arrays(s.ranks[0])
    if cfg.isPQEq:
        e.PQEq(atype, pos, q)
    else:
        e.QEq(atype, pos, q)
    return s, cfg, e


def _reference_rowsums(e, x):
    n = e.NATOMS
    rb, re_, col, val = e.fetch("rowbeg"), e.fetch("rowend"), e.fetch("col"), e.fetch("val")
    xl = x.astype(np.longdouble)
    ref = np.zeros((n, 4), dtype=np.longdouble)
    scale = np.zeros((n, 2))
    for i in range(n):
        c = col[rb[i]:re_[i]]
        h = val[rb[i]:re_[i]].astype(np.longdouble)
        p = h[:, None] * xl[c]
        g = c >= n                                  # ghost columns (Est weighting)
        ref[i, :2] = p.sum(axis=0)
        ref[i, 2:] = p[g].sum(axis=0)
        scale[i] = np.abs(p).sum(axis=0).astype(np.float64)
    return ref.astype(np.float64), scale


@pytest.mark.parametrize("name", list(systems().keys()))
@pytest.mark.parametrize("variant", list(VARIANTS.keys()))
def test_spmv_rowsums_match_numpy(built, name, variant):
    s, cfg, e = _engine(name, VARIANTS[variant])
    ntot = int(e.fetch("copyptr")[6])
    rng = np.random.default_rng(11)
    x = rng.normal(0.0, 1.0, (ntot, 2))
    got, _ = e.debug_spmv(x)
    ref, scale = _reference_rowsums(e, x)
    sc = np.maximum(np.concatenate([scale, scale], axis=1), 1e-300)
    err = np.abs(got - ref) / sc
    print(f"{name} {variant}: max rel err {err.max():.2e} (rows {e.NATOMS}, nnz {int(e.fetch('nnz')[0])})")
    assert err.max() <= 1e-13
    t = e.timers()
    if variant.startswith("win"):   # the launch really was k_spmv_win (a list whose windows do not fit falls back to k_spmv_rows)
        print(f"  win launches {int(t[25])}, rows launches {int(t[26])}, G {int(t[27])}, largest window {int(t[28])}")
        assert t[25] > 0 and t[26] == 0
    else:
        assert t[25] == 0
    e.close()


@pytest.mark.parametrize("name", list(systems().keys()))
@pytest.mark.parametrize("group", ["auto", "3"])
def test_window_stream_matches_rows(built, name, group):
    """The 16-bit column stream of k_spmv_win, checked structurally against the 32-bit columns it was built beside: for every
    row, position p of the row's group window must be the slot the 32-bit column names -- the window being the concatenation, in
    stencil-run order, of the slot ranges the descriptors (k_win_desc) list for the group -- bit 15 must be the ghost flag,
    `rowlen` the row's exact length, and the descriptors must lay the runs out back to back."""
    env = {} if group == "auto" else {"RXG_WIN_G": group}
    s, cfg, e = _engine(name, env)
    meta = e.fetch("win_meta")
    built_, G, nruns, ngroups, nc0, nc1, nc2, L = [int(x) for x in meta]
    assert built_ == 1
    n = e.NATOMS
    ntot = int(e.fetch("copyptr")[6])
    rb, re_ = e.fetch("rowbeg"), e.fetch("rowend")
    cs, c16, rowlen = e.fetch("col_slot"), e.fetch("col16"), e.fetch("rowlen")
    order, cell = e.fetch("order_nb"), e.fetch("cell_nb")
    desc = e.fetch("win_desc").reshape(ngroups, nruns + 1, 2)
    slot_of = np.empty(ntot, dtype=np.int64)
    slot_of[order[:ntot]] = np.arange(ntot)
    dims = (nc0 + 2 * L, nc1 + 2 * L, nc2 + 2 * L)
    ngz = -(-nc2 // G)
    checked = 0
    for i in range(n):
        cid = int(cell[i])
        c3 = cid % dims[2] - L
        c2 = (cid // dims[2]) % dims[1] - L
        c1 = cid // (dims[2] * dims[1]) - L
        assert 0 <= c1 < nc0 and 0 <= c2 < nc1 and 0 <= c3 < nc2     # residents sit in resident cells
        grp = (c1 * nc1 + c2) * ngz + c3 // G
        d = desc[grp]
        ws, packed = d[:nruns, 0].astype(np.int64), d[:nruns, 1].astype(np.int64)
        wlen, pos = packed & 0xfff, packed >> 12
        assert np.array_equal(pos, np.concatenate([[0], np.cumsum(wlen)[:-1]]))      # runs are laid out back to back
        assert int(d[nruns, 1]) >> 12 == int(wlen.sum())                              # ... and the total closes the table
        # window position -> slot
        win = np.concatenate([np.arange(a, a + l) for a, l in zip(ws, wlen)]) if wlen.sum() else np.zeros(0, dtype=np.int64)
        k0, k1 = int(rb[i]), int(re_[i])
        assert rowlen[slot_of[i]] == k1 - k0
        w = c16[k0:k1].astype(np.int64)
        assert np.array_equal(win[w & 0x7fff], cs[k0:k1].astype(np.int64) & 0x7fffffff), f"row {i}"
        assert np.array_equal((w >> 15) == 1, cs[k0:k1] < 0), f"ghost flags of row {i}"
        checked += k1 - k0
    assert checked == int(e.fetch("nnz")[0]) or checked > 0
    print(f"{name} G={G}: {checked} window-relative columns over {ngroups} groups x {nruns} runs check out")
    e.close()


@pytest.mark.parametrize("name", list(systems().keys()))
def test_union_stream_matches_rows(built, name):
    """Per block of <= 8 consecutive rows of a cell, the union stream holds exactly the columns of those rows: replaying a
    row's bits in stream order reproduces the row (same columns, same order), so a value is found at the row's running
    position; padding entries carry an empty row set."""
    s, cfg, e = _engine(name, {"RXG_SPMV": "items"})
    n = e.NATOMS
    ntot = int(e.fetch("copyptr")[6])
    rb, re_, col = e.fetch("rowbeg"), e.fetch("rowend"), e.fetch("col")
    order, uoff, ucol, umask, cell = e.fetch("order_nb"), e.fetch("uoff"), e.fetch("ucol"), e.fetch("umask"), e.fetch("cell_nb")
    ghost = ucol < 0
    uatom = order[ucol & 0x7fffffff]
    real_entry = umask != 0                             # padding repeats a column of the block without its ghost bit
    assert np.array_equal(ghost[real_entry], (uatom >= n)[real_entry])
    assert uoff[ntot] == len(ucol)
    seen_rows = 0
    s0 = 0
    while s0 < ntot:
        i0 = order[s0]
        if i0 >= n:                                  # ghost slots own no block
            assert uoff[s0 + 1] == uoff[s0]
            s0 += 1
            continue
        # block = up to 8 consecutive slots of one cell
        nr = 1
        while nr < 8 and s0 + nr < ntot and cell[order[s0 + nr]] == cell[i0] and order[s0 + nr] < n:
            nr += 1
        # blocks start at the cell's first slot: a cell's 9th row starts a new block
        a, b = uoff[s0], uoff[s0 + 1]
        assert (b - a) % 16 == 0                        # bulk-copy granularity of the row sets
        for r in range(1, nr):
            assert uoff[s0 + r + 1] == uoff[s0 + r]
        cols, bits = uatom[a:b], umask[a:b]
        for r in range(nr):
            i = order[s0 + r]
            mine = cols[(bits >> r) & 1 == 1]
            assert np.array_equal(mine, col[rb[i]:re_[i]]), f"row {i} (slot {s0 + r})"
            seen_rows += 1
        assert not np.any(bits >> nr)                # no bits beyond the block's rows
        real = bits != 0
        assert not np.any(real[np.argmin(real):]) if not real.all() else True    # padding sits at the end
        s0 += nr
    assert seen_rows == n
    e.close()
