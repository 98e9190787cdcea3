"""world_size-2 (and 4) run of the host-side N>1 logic on CPU with the gloo backend:
decomposition (eulerb200_decompose), the exchange plan (eulerb200_exchange_plan) driven over
torch.distributed point-to-point, halo layers in the reference's wire layout -- with the
CPU oracle standing in for the kernels.  The assembled result must equal the single-rank
oracle BIT FOR BIT (decomposition invariance)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P, N, D, R = 0, 1, 2, 3
NO = -1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port_no, n, nchem, bcs, outdir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    import oracle
    from __graft_entry__ import load_package
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port_no)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pkg = load_package()
    port = oracle.Port()
    d = (0.1, 0.2, 0.3)
    w = oracle.random_state(n, nchem, seed=77)          # every rank builds the same global state
    W3 = [w[f].reshape(n[2], n[1], n[0]) for f in range(5)]
    W3.append(w[5].reshape(n[2], n[1], n[0], nchem) if nchem else None)

    u = pkg.EulerData(nchem=nchem)
    u.nx, u.ny, u.nz = n
    u.xlbc, u.xrbc, u.ylbc, u.yrbc, u.zlbc, u.zrbc = bcs
    rc, dims, coords, ext, nbr = pkg.dims_and_extents(world, rank, n, bcs)
    assert rc == 0
    u.myid, u.nprocs = rank, world
    u.nxl, u.nyl, u.nzl = ext[1] - ext[0] + 1, ext[3] - ext[2] + 1, ext[5] - ext[4] + 1
    u.dx, u.dy, u.dz = d
    u.ipW, u.ipE, u.ipS, u.ipN, u.ipB, u.ipF = nbr
    sl = (slice(ext[4], ext[5] + 1), slice(ext[2], ext[3] + 1), slice(ext[0], ext[1] + 1))
    parts = [np.ascontiguousarray(a[sl]).ravel() if a is not None else None for a in W3]
    nl = (u.nxl, u.nyl, u.nzl)
    cfg = port.cfg(nl, nchem, d, 1.4, bcs)

    # the exchange, operation by operation in the order the library issues them
    recv = [None] * 6
    reqs, keep = [], []
    for kind, face, peer in u.exchange_plan():
        if kind == "send":
            t = torch.from_numpy(port.pack_send(cfg, parts, face))
            keep.append(t)
            reqs.append(dist.isend(t, dst=peer))
        else:
            t = torch.empty(port.face_len(cfg, face), dtype=torch.float64)
            recv[face] = t
            reqs.append(dist.irecv(t, src=peer))
    for r in reqs:
        r.wait()
    ext_bc = [(-1 if recv[f] is not None else (bcs[f])) for f in range(6)]
    cfg2 = port.cfg(nl, nchem, d, 1.4, ext_bc)
    ret, wdot, _ = port.feuler(cfg2, parts, ext=[r.numpy() if r is not None else None for r in recv])
    assert ret == 0
    np.savez(os.path.join(outdir, "rank%d.npz" % rank), ext=np.array(ext),
             **{"wdot%d" % f: wdot[f] for f in range(6) if wdot[f] is not None})
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n,nchem,bcs", [
    (2, (12, 8, 7), 2, [P, P, R, R, N, N]),
    (2, (3, 14, 9), 0, [N] * 6),              # thin x: the split goes to y
    (4, (12, 12, 6), 2, [P] * 6),
])
def test_decomposed_exchange_and_rhs_equal_single_rank(tmp_path, world, n, nchem, bcs):
    import torch.multiprocessing as mp
    sys.path.insert(0, ROOT)
    import oracle
    port_no = _free_port()
    mp.spawn(_worker, args=(world, port_no, n, nchem, bcs, str(tmp_path)), nprocs=world, join=True)
    port = oracle.Port()
    w = oracle.random_state(n, nchem, seed=77)
    ret, ref, _ = port.feuler(port.cfg(n, nchem, (0.1, 0.2, 0.3), 1.4, bcs), w)
    assert ret == 0
    R3 = [ref[f].reshape(n[2], n[1], n[0]) for f in range(5)]
    if nchem:
        R3.append(ref[5].reshape(n[2], n[1], n[0], nchem))
    for rank in range(world):
        z = np.load(os.path.join(str(tmp_path), "rank%d.npz" % rank))
        ext = z["ext"]
        sl = (slice(ext[4], ext[5] + 1), slice(ext[2], ext[3] + 1), slice(ext[0], ext[1] + 1))
        for f, a in enumerate(R3):
            assert np.array_equal(np.ascontiguousarray(a[sl]).ravel(), z["wdot%d" % f])
