"""Host-side logic of the N>1 path on CPU: world_size-2 gloo processes.

What runs here is everything around the CUDA library that a multi-rank run needs: the consistent partition of the
atoms over `vprocs`, the neighbour topology, the rendezvous of the 128-byte communicator id over torch.distributed, and
the reduction of per-rank observables (PRINTE's MPI_ALLREDUCE, src/main.F90:242) -- checked against the oracle, which
simulates the same two ranks in-process."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "tests", "golden", "inputs", "init.rdx")


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    import sys
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from rxmd_b200.host.system import build_system
    from rxmd_b200.host.engine import broadcast_bytes, rank_of_vid
    from oracle.pyoracle import Oracle
    s = build_system(os.path.join(G, "input.xyz"), os.path.join(G, "ffield"), mc=(2, 1, 1), vprocs=(2, 1, 1), displace_sigma=0.02)
    # 1. id rendezvous (the bytes MPI_Bcast would carry)
    payload = bytes(range(128)) if rank == 0 else bytes(128)
    got = broadcast_bytes(dist, payload)
    ok_id = got == bytes(range(128))
    # 2. partition: my atoms are mine and only mine
    b = s.boxes[rank].struct
    mine = len(s.ranks[rank]["atype"])
    t = torch.tensor([mine], dtype=torch.int64)
    dist.all_reduce(t)
    ok_part = int(t.item()) == s.natoms
    tn = list(b.target_node)
    ok_topo = tn[0] == tn[1] == 1 - rank and tn[2:] == [rank] * 4 and rank_of_vid(list(b.vID), (2, 1, 1)) == rank
    # 3. per-rank energies reduce to the global ones
    o = Oracle(s, s.config(nbuffer=30000))
    o.qeq(); o.force()
    pe_mine = torch.tensor(o.f64("PE", rank)[1:], dtype=torch.float64)
    dist.all_reduce(pe_mine)
    pe_all = o.observe()[0][1:]
    ok_pe = bool(np.allclose(pe_mine.numpy(), pe_all, rtol=1e-13, atol=1e-10))
    out[rank] = (ok_id, ok_part, ok_topo, ok_pe)
    o.close()
    dist.destroy_process_group()


def test_two_ranks_gloo(built):
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    assert dict(out) == {0: (True, True, True, True), 1: (True, True, True, True)}, dict(out)
