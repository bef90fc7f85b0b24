"""Host-side restatements (ffield parser, geninit, rxff.bin, stencil) against facts recorded from the reference inputs."""
import os

import numpy as np

from rxmd_b200.host.ffield import read_ffield
from rxmd_b200.host import geninit, setup as S
from rxmd_b200.host.system import build_system

INP = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "inputs")


def test_ffield_nitramine():
    ff = read_ffield(os.path.join(INP, "init.rdx", "ffield"))
    assert ff.atmname[1:5] == ["C", "H", "O", "N"]              # H must be type 2 for Ehb (src/pot.F90:595)
    assert ff.nso >= 4 and ff.nboty >= 10
    assert ff.inxn2[1, 2] == ff.inxn2[2, 1] != 0
    assert np.all(ff.eta[1:5] > 0)
    # pair cut-offs, SURVEY App. C (rule src/init.F90:387-395)
    npt = np.zeros(ff.nso + 1, dtype=int); npt[1:5] = [24, 48, 48, 48]
    rc, rc2, maxrc = S.cutoff_length(ff, npt)
    want = {(1, 1): 2.62, (1, 2): 2.15, (2, 2): 1.89, (1, 3): 3.15, (3, 3): 2.44, (1, 4): 2.84, (3, 4): 2.65, (4, 4): 3.16,
            (2, 3): 2.27, (2, 4): 2.06}
    for (a, b), r in want.items():
        assert abs(rc[ff.inxn2[a, b]] - r) < 1e-9, (a, b, rc[ff.inxn2[a, b]])
    assert abs(maxrc - 3.16) < 1e-9


def test_water_hbond_quirk():
    """conf/init.water/ffield lists H first: type 2 is O, so the reference's Ehb (jty==2) never fires."""
    ff = read_ffield(os.path.join(INP, "init.water", "ffield"))
    assert ff.atmname[1] == "H" and ff.atmname[2] == "O"


def test_replicate_and_rxff_roundtrip(tmp_path):
    ff = read_ffield(os.path.join(INP, "init.rdx", "ffield"))
    t0, p0, lat = geninit.read_xyz(os.path.join(INP, "init.rdx", "input.xyz"), ff.atmname)
    gen = geninit.replicate(t0, p0, lat, (2, 1, 2), (2, 1, 1))
    assert gen["natoms"] == 168 * 4 and sum(len(r["atype"]) for r in gen["ranks"]) == 168 * 4
    for r in gen["ranks"]:
        assert r["pos_local"].min() >= 0 and r["pos_local"][:, 0].max() < 0.5 + 1e-12
    gids = np.concatenate([np.rint((r["atype"] - np.rint(r["atype"])) * 1e13) for r in gen["ranks"]]).astype(int)
    assert sorted(gids) == list(range(1, 168 * 4 + 1))
    p = tmp_path / "rxff.bin"
    geninit.write_rxff_bin(str(p), gen, (2, 1, 1))
    back = geninit.read_rxff_bin(str(p))
    assert back["vprocs"] == (2, 1, 1) and back["natoms"] == gen["natoms"]
    for a, b in zip(gen["ranks"], back["ranks"]):
        assert np.array_equal(a["pos_local"], b["pos_local"]) and np.array_equal(a["atype"], b["atype"])
    assert os.path.getsize(p) == 4 * (4 + 2 + 1) + 48 + 80 * gen["natoms"]     # src/fileio.F90:477-505


def test_stencil_sizes():
    """305 stencil cells for the 168-atom box, 477 for large boxes."""
    nbcc, nbl, mesh = S.nonbonding_mesh(13.18, 11.57, 10.71, (1, 1, 1), 10.0)
    assert len(mesh) == 305
    nbcc, nbl, mesh = S.nonbonding_mesh(13.18 * 18, 11.57 * 18, 10.71 * 18, (1, 1, 1), 10.0)
    assert len(mesh) == 477 and list(nbcc) == [79, 69, 64]


def test_build_system_partitions_consistently():
    g = os.path.join(INP, "init.rdx")
    s = build_system(os.path.join(g, "input.xyz"), os.path.join(g, "ffield"), mc=(2, 2, 1), vprocs=(2, 2, 1), displace_sigma=0.02)
    assert sum(len(r["atype"]) for r in s.ranks) == s.natoms == 168 * 4
    for r, b in zip(s.ranks, s.boxes):
        Hi = np.asarray(b.Hi)
        rn = (Hi @ r["pos"]).T - np.array(list(b.struct.OBOX))
        assert rn.min() > -1e-12 and np.all(rn.max(axis=0) < np.array(list(b.struct.LBOX)) + 1e-12)


def test_per_rank_generation_equals_global_generation(tmp_path):
    """a rank generates (and reads) its own sub-domain only.  `only_rank=r` must give exactly the atoms the
    all-rank build gives rank r, the synthetic displacements must not depend on the decomposition, and the rxff.bin slice
    reader must return rank r's records alone (64-bit offsets, src/fileio.F90:499-505)."""
    g = os.path.join(INP, "init.rdx.lg")
    xyz, ffp = os.path.join(g, "input.xyz"), os.path.join(g, "ffield")
    full = build_system(xyz, ffp, mc=(4, 4, 2), vprocs=(2, 2, 1), isLG=True, displace_sigma=0.02)
    one = build_system(xyz, ffp, mc=(4, 4, 2), vprocs=(1, 1, 1), isLG=True, displace_sigma=0.02)
    for r in range(4):
        part = build_system(xyz, ffp, mc=(4, 4, 2), vprocs=(2, 2, 1), isLG=True, displace_sigma=0.02, only_rank=r)
        assert [x is None for x in part.ranks] == [k != r for k in range(4)]
        assert np.array_equal(part.ranks[r]["atype"], full.ranks[r]["atype"])
        assert np.array_equal(part.ranks[r]["pos"], full.ranks[r]["pos"])
        assert part.natoms == full.natoms == 168 * 32
    # same geometry whatever the decomposition: match atoms by global id
    gid = lambda a: np.rint((a - np.rint(a)) * 1e13).astype(np.int64)
    a4 = np.concatenate([r["atype"] for r in full.ranks]); p4 = np.concatenate([r["pos"] for r in full.ranks], axis=1)
    o4, o1 = np.argsort(gid(a4)), np.argsort(gid(one.ranks[0]["atype"]))
    assert np.array_equal(gid(a4)[o4], gid(one.ranks[0]["atype"])[o1])
    assert np.abs(p4[:, o4] - one.ranks[0]["pos"][:, o1]).max() < 1e-9
    # displacements are Gaussian with the requested width
    und = build_system(xyz, ffp, mc=(4, 4, 2), isLG=True)
    d = one.ranks[0]["pos"] - und.ranks[0]["pos"]
    d = d[:, np.abs(d).max(axis=0) < 1.0]                            # skip atoms wrapped through a face
    assert abs(d.std() - 0.02) < 1e-3 and abs(d.mean()) < 1e-3
    # rxff.bin: per-rank slice
    ff = read_ffield(ffp, isLG=True)
    t0, p0, lat = geninit.read_xyz(xyz, ff.atmname)
    gen = geninit.replicate(t0, p0, lat, (4, 4, 2), (2, 2, 1))
    p = tmp_path / "rxff.bin"
    geninit.write_rxff_bin(str(p), gen, (2, 2, 1))
    sl = geninit.read_rxff_bin(str(p), only_rank=2)
    assert [x is None for x in sl["ranks"]] == [True, True, False, True]
    assert np.array_equal(sl["ranks"][2]["atype"], gen["ranks"][2]["atype"]) and np.array_equal(sl["ranks"][2]["pos_local"], gen["ranks"][2]["pos_local"])
    assert sl["natoms_per_rank"] == tuple(len(r["atype"]) for r in gen["ranks"])
