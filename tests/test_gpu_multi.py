"""Multi-GPU tier (needs >= 2 B200s; skipped otherwise): the decomposed CUDA RHS with NCCL
halo exchange and interior/boundary overlap must equal the single-rank oracle."""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P, N, D, R = 0, 1, 2, 3


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port_no, n, nchem, bcs, outdir, transport, overlap="1"):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    import oracle
    from __graft_entry__ import load_package
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port_no)
    os.environ["EULERB200_HALO"] = transport
    os.environ["EULERB200_OVERLAP"] = overlap[0]
    os.environ["EULERB200_SHELLS"] = "1" if overlap.endswith("t") else "0"      # "1t": tile-thick boundary shells
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    pkg = load_package()
    u = pkg.EulerData(nchem=nchem)
    u.nx, u.ny, u.nz = n
    u.xlbc, u.xrbc, u.ylbc, u.yrbc, u.zlbc, u.zrbc = bcs
    u.forcing = [0, 0, -0.1, 0, 0]
    assert u.SetupDecomp(myid=rank, nprocs=world, device=rank) == 0
    assert u.halo_transport == transport
    w = oracle.random_state(n, nchem, seed=31)
    W3 = [w[f].reshape(n[2], n[1], n[0]) for f in range(5)] + ([w[5].reshape(n[2], n[1], n[0], nchem)] if nchem else [])
    sl = (slice(u.ks, u.ke + 1), slice(u.js, u.je + 1), slice(u.is_, u.ie + 1))
    wl = pkg.ManyVector([torch.from_numpy(np.ascontiguousarray(a[sl]).ravel()).cuda() for a in W3])
    wdot = pkg.ManyVector.new(u)
    for _ in range(3):       # repeatedly: slabs (both parities), flags and events must be reusable
        assert pkg.fEuler(0.0, wl, wdot, u) == 0, u.last_error()
    u.cfl = 0.4
    ret, dt = pkg.stability(wl, 0.0, u)
    assert ret == 0
    np.savez(os.path.join(outdir, "rank%d.npz" % rank), ext=np.array([u.is_, u.ie, u.js, u.je, u.ks, u.ke]),
             dt=dt, **{"wdot%d" % f: s.cpu().numpy() for f, s in enumerate(wdot.sub)})
    dist.barrier()
    u.FreeData()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n,nchem,bcs,transport,overlap", [
    (2, (40, 24, 20), 2, [P, P, R, R, N, N], "p2p", "1"),      # periodic in x over 2 ranks: both x-faces go to the same peer
    (2, (40, 24, 20), 2, [P, P, R, R, N, N], "nccl", "1"),
    (2, (40, 24, 20), 2, [P, P, R, R, N, N], "nccl", "0"),
    (2, (80, 30, 20), 10, [R] * 6, "nccl", "0"),               # full-width (32-column) tiles next to a rank seam
    (2, (80, 30, 20), 10, [R] * 6, "p2p", "2"),
    (2, (40, 24, 20), 2, [P, P, R, R, N, N], "nccl", "2"),
    (2, (272, 50, 12), 2, [R] * 6, "nccl", "1t"),              # 136 columns per rank: tile-thick x-shell (32 columns)
    (2, (272, 50, 12), 2, [P, P, R, R, N, N], "p2p", "2t"),    # ... on both sides (periodic over 2 ranks)
    (2, (272, 50, 12), 2, [P, P, R, R, N, N], "nccl", "1"),
    (4, (272, 100, 8), 2, [P] * 6, "nccl", "2t"),              # tile-thick shells in x and y
    (2, (3, 40, 36), 0, [N] * 6, "p2p", "1"),
    (4, (24, 28, 20), 2, [P] * 6, "p2p", "0"),
    (4, (24, 28, 20), 2, [P] * 6, "nccl", "2"),
    (8, (24, 24, 24), 10, [R] * 6, "p2p", "2"),
    (8, (24, 24, 24), 10, [R] * 6, "nccl", "0"),
    (8, (3, 64, 48), 0, [N] * 6, "nccl", "1"),
])
def test_decomposed_cuda_rhs_equals_single_rank_oracle(tmp_path, world, n, nchem, bcs, transport, overlap):
    """transport: "p2p" = pack kernels store into the neighbour's ghost slab over NVLink (CUDA IPC)
    and publish a sequence number; "nccl" = pack + grouped ncclSend/ncclRecv on a side stream.
    overlap (EULERB200_OVERLAP): "1" = interior launch behind the exchange, then the boundary shells; "2" = the
    shells on high-priority streams as soon as the halo is in, next to the interior launch; "0" = exchange first
    (behind the pre-pass), then one launch over the whole box; a trailing "t" = tile-thick boundary shells
    (EULERB200_SHELLS=1)."""
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    import torch.multiprocessing as mp
    sys.path.insert(0, ROOT)
    import oracle
    mp.spawn(_worker, args=(world, _free_port(), n, nchem, bcs, str(tmp_path), transport, overlap), nprocs=world, join=True)
    port = oracle.Port()
    w = oracle.random_state(n, nchem, seed=31)
    d = (1.0 / n[0], 1.0 / n[1], 1.0 / n[2])
    cfg = port.cfg(n, nchem, d, 1.4, bcs, forcing=[0, 0, -0.1, 0, 0])
    ret, ref, _ = port.feuler(cfg, w)
    assert ret == 0
    dt_want = port.dt_stab(cfg, 0.4, port.max_wavespeed(cfg, w))
    R3 = [ref[f].reshape(n[2], n[1], n[0]) for f in range(5)] + ([ref[5].reshape(n[2], n[1], n[0], nchem)] if nchem else [])
    for rank in range(world):
        z = np.load(os.path.join(str(tmp_path), "rank%d.npz" % rank))
        e = z["ext"]
        sl = (slice(e[4], e[5] + 1), slice(e[2], e[3] + 1), slice(e[0], e[1] + 1))
        assert float(z["dt"]) == pytest.approx(dt_want, rel=1e-14)
        for f, a in enumerate(R3):
            want = np.ascontiguousarray(a[sl]).ravel()
            assert np.abs(z["wdot%d" % f] - want).max() <= 1e-12 * np.abs(a).max()


def _halo_worker(rank, world, port_no, nvar, outdir, transport):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from __graft_entry__ import load_package
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port_no)
    os.environ["EULERB200_HALO"] = transport
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    pkg = load_package()
    n = (12, 12, 12)
    u = pkg.EulerData(nchem=nvar - 5)
    u.nx, u.ny, u.nz = n
    assert u.SetupDecomp(myid=rank, nprocs=world, device=rank) == 0          # all-periodic by default
    idx = np.arange(n[0] * n[1] * n[2])
    i, j, k = idx % n[0], (idx // n[0]) % n[1], idx // (n[0] * n[1])
    enc = lambda v: (0.001 * v + 1e-6 * i + 1e-9 * j + 1e-12 * k).reshape(n[2], n[1], n[0])
    sl = (slice(u.ks, u.ke + 1), slice(u.js, u.je + 1), slice(u.is_, u.ie + 1))
    subs = [torch.from_numpy(np.ascontiguousarray(enc(v)[sl]).ravel()).cuda() for v in range(5)]
    if nvar > 5:
        subs.append(torch.from_numpy(np.ascontiguousarray(np.stack([enc(v)[sl] for v in range(5, nvar)], axis=-1)).ravel()).cuda())
    w = pkg.ManyVector(subs)
    for rep in range(2):
        assert u.ExchangeStart(w) == 0 and u.ExchangeEnd() == 0
    torch.cuda.synchronize()
    out = {"ext": np.array([u.is_, u.ie, u.js, u.je, u.ks, u.ke]), "nbr": np.array(u.nbrs)}
    for f in range(6):
        out["recv%d" % f] = u.recv_buffer(w, f).cpu().numpy()
    np.savez(os.path.join(outdir, "halo%d.npz" % rank), **out)
    dist.barrier()
    u.FreeData()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,nvar,transport", [(2, 5, "nccl"), (2, 7, "p2p"), (8, 7, "nccl"), (8, 5, "p2p")])
def test_exchange_equals_the_references_own_receive_buffers(tmp_path, world, nvar, transport):
    """communication_test_main.cpp on GPUs: after ExchangeStart/ExchangeEnd the six ghost slabs of
    every rank must equal, entry for entry, what the REFERENCE's ExchangeStart/ExchangeEnd delivered
    on the same 12^3 periodic grid with (field, i, j, k)-encoded values (tests/golden/exchange_*.npz,
    produced by the unmodified reference on 2 and 8 virtual MPI ranks)."""
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    import torch.multiprocessing as mp
    mp.spawn(_halo_worker, args=(world, _free_port(), nvar, str(tmp_path), transport), nprocs=world, join=True)
    gold = np.load(os.path.join(ROOT, "tests", "golden", "exchange_nvar%d.npz" % nvar))
    for rank in range(world):
        z = np.load(os.path.join(str(tmp_path), "halo%d.npz" % rank))
        assert list(z["ext"]) == list(gold["p%d_r%d_ext" % (world, rank)])
        assert [int(x) for x in z["nbr"]] == [int(x) for x in gold["p%d_r%d_nbr" % (world, rank)]]
        for f in range(6):
            assert np.array_equal(z["recv%d" % f], gold["p%d_r%d_recv%d" % (world, rank, f)]), (rank, f)


# ---- full runs: diagnostics must not depend on the decomposition (euler3D_main.cpp:378-417) -------------

def _run_worker(rank, world, port_no, problem, n, tf, opts, outdir):
    sys.path.insert(0, ROOT)
    import json
    import torch
    import torch.distributed as dist
    from __graft_entry__ import load_package
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port_no)
    torch.cuda.set_device(rank)
    if world > 1:
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    pkg = load_package()
    u = pkg.EulerData(nchem=0)
    u.nx, u.ny, u.nz = n
    pkg.problems.configure(problem, u)
    assert u.SetupDecomp(myid=rank, nprocs=world, device=rank) == 0
    w = pkg.ManyVector.new(u)
    assert pkg.problems.initial_conditions(problem, 0.0, w, u) == 0
    step = pkg.driver.ERKStep(pkg.driver.TorchVecOps(pkg, u), 0.0, w, pkg.driver.ARKODEParameters(**opts))
    cons = pkg.problems.Conservation()
    rec = {"cons0": cons(0.0, w, u, quiet=True), "outputs": []}
    for iout in range(2):
        ret, t = step.evolve(tf * (iout + 1) / 2)
        assert ret == 0
        rec["outputs"].append({"t": t, "diag": pkg.problems.output_diagnostics(problem, t, step.w, u, quiet=True),
                               "stats": pkg.problems.print_stats(t, step.w, u, step.nst, quiet=True),
                               "cons": cons(t, step.w, u, quiet=True)})
    rec["solver"] = step.stats()
    rec["dims"] = [u.npx, u.npy, u.npz]
    if rank == 0:
        with open(os.path.join(outdir, "run_%s_%d.json" % (problem, world)), "w") as f:
            json.dump(rec, f)
    if world > 1:
        dist.barrier()
    u.FreeData()
    if world > 1:
        dist.destroy_process_group()


@pytest.mark.parametrize("problem,n,tf,opts", [
    ("sod_x", (96, 12, 12), 0.02, dict(order=4, fixedstep=1, hmax=0.001)),
    ("rayleigh_taylor", (24, 72, 12), 0.02, dict(order=4, fixedstep=1, hmax=0.002)),
])
def test_full_runs_print_the_same_diagnostics_on_1_2_4_8_gpus(tmp_path, problem, n, tf, opts):
    """Fixed-step Sod and Rayleigh-Taylor (with its forcing) runs through the explicit driver on 1, 2, 4 and 8
    GPUs (as many as the box has): errI / errR against the analytic solution (Sod), the statistics table,
    the conservation totals and the solver counters must not depend on the decomposition -- every face
    flux is computed from the same six values whichever rank owns it; what differs is
    the order of the cross-rank sums and last-bit rounding in cells next to a rank seam."""
    import json
    import torch
    import torch.multiprocessing as mp
    worlds = [wd for wd in (1, 2, 4, 8) if wd <= torch.cuda.device_count()]
    if len(worlds) < 2:
        pytest.skip("needs >= 2 GPUs")
    for wd in worlds:
        mp.spawn(_run_worker, args=(wd, _free_port(), problem, n, tf, opts, str(tmp_path)), nprocs=wd, join=True)
    recs = {wd: json.load(open(os.path.join(str(tmp_path), "run_%s_%d.json" % (problem, wd)))) for wd in worlds}
    base = recs[1]
    close = lambda a, b: abs(a - b) <= 1e-12 * max(abs(a), abs(b), 1e-300)
    for wd in worlds[1:]:
        r = recs[wd]
        assert r["dims"][0] * r["dims"][1] * r["dims"][2] == wd
        assert r["solver"] == base["solver"]
        assert close(r["cons0"]["mass"], base["cons0"]["mass"]) and close(r["cons0"]["energy"], base["cons0"]["energy"])
        for a, b in zip(r["outputs"], base["outputs"]):
            assert a["t"] == b["t"]
            assert all(close(x, y) for x, y in zip(a["stats"], b["stats"]))
            assert close(a["cons"]["mass"], b["cons"]["mass"]) and close(a["cons"]["energy"], b["cons"]["energy"])
            if b["diag"] is not None:
                near = lambda x, y: abs(x - y) <= 1e-9 * max(abs(x), abs(y)) + 1e-14   # cells next to a rank seam take the
                assert all(near(x, y) for x, y in zip(a["diag"]["errI"], b["diag"]["errI"]))   # boundary-tile code path: last-bit
                assert all(near(x, y) for x, y in zip(a["diag"]["errR"], b["diag"]["errR"]))   # differences in the state
