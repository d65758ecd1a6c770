"""CPU, world_size 2, gloo: the band-sharding host logic (state blocks + the rho / scalar all-reduces).  Each rank builds
the density and E_nl of ITS state block (with the oracle standing in for the GPU kernels, which need a device) and the
all-reduced result must equal the single-rank result."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from qball_b200 import parallel as PAR


def test_state_block_partition():
    for nst, world in [(768, 8), (10118, 8), (5, 2), (3, 4), (7, 8)]:
        blocks = [PAR.state_block(nst, r, world) for r in range(world)]
        assert sum(b[1] for b in blocks) == nst
        nb = -(-nst // world)
        for r, (f, n) in enumerate(blocks):
            assert f == min(r * nb, nst) and 0 <= n <= nb   # SlaterDet.cc:228-231: nb = ceil(nst/npcol)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    here = os.path.dirname(os.path.abspath(__file__))
    for p in (os.path.dirname(here), os.path.join(os.path.dirname(here), "oracle"), here):
        if p not in sys.path:
            sys.path.insert(0, p)
    import port as P
    import refdrive as R
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cell, ecut, nst = (9, 0, 0, 0, 9, 0, 0, 0, 10), 4.0, 5
    b = P.make_basis(cell, ecut, (0.25, 0, 0))
    grid = P.density_grid(cell, ecut)
    occ = R.synth_occ(nst, 4)
    first, n = PAR.state_block(nst, rank, world)
    c = R.synth_coefficients(b["kpg2"], ecut, n, b["ngw"], b["is_real"], seed=3, first_state=first)
    ft = P.FT(b, *grid)
    rho = ft.compute_density(c, occ[first:first + n] / b["omega"], np.zeros(ft.N))
    rng = np.random.default_rng(0)
    sp = [dict(na=2, npr=4, lproj=np.array([0, 1, 1, 1], dtype=np.int32), wt=np.array([1.0, -0.5, -0.5, -0.5]),
               twnl=rng.standard_normal((4, b["ngw"])), tau=rng.uniform(0, 9, (2, 3)))]
    enl, _ = P.nl_energy(b, c, occ[first:first + n], sp)
    nel_local = float(rho.sum() * b["omega"] / ft.N)          # before the in-place all-reduce (ChargeDensity.cc:525-528)
    rho_np = rho.copy()                                        # numpy HOST array: must be summed over ranks too
    assert PAR.allreduce_density(rho_np) is rho_np
    rho_t = PAR.allreduce_density(torch.from_numpy(rho))
    assert np.array_equal(rho_np, rho_t.numpy())
    enl_sum, nel = PAR.allreduce_scalars([enl, nel_local])
    if rank == 0:
        call = R.synth_coefficients(b["kpg2"], ecut, nst, b["ngw"], b["is_real"], seed=3)
        rho_all = ft.compute_density(call, occ / b["omega"], np.zeros(ft.N))
        enl_all, _ = P.nl_energy(b, call, occ, sp)
        q.put((float(np.abs(rho_t.numpy() - rho_all).max() / np.abs(rho_all).max()), abs(enl_sum - enl_all) / abs(enl_all),
               abs(nel - rho_all.sum() * b["omega"] / ft.N)))
    dist.barrier()
    dist.destroy_process_group()


def test_band_sharded_density_and_enl_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 500
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0] < 1e-12 and res[1] < 1e-12 and res[2] < 1e-10, res


def _worker_la(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    here = os.path.dirname(os.path.abspath(__file__))
    for p in (os.path.dirname(here), os.path.join(os.path.dirname(here), "oracle"), here):
        if p not in sys.path:
            sys.path.insert(0, p)
    import port as P
    import refdrive as R
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cell, ecut, nst = (9, 0, 0, 0, 9, 0, 0, 0, 10), 4.0, 5           # 5 states on 2 ranks: blocks of 3 and 2
    b = P.make_basis(cell, ecut, (0.25, 0, 0))
    first, n = PAR.state_block(nst, rank, world)
    c = R.synth_coefficients(b["kpg2"], ecut, n, b["ngw"], b["is_real"], seed=3, first_state=first)
    hc = R.synth_coefficients(b["kpg2"], ecut, n, b["ngw"], b["is_real"], seed=4, first_state=first)
    call = PAR.allgather_states(torch.from_numpy(c), nst).numpy()    # the exchange step
    res, a = P.residual(np.ascontiguousarray(call), hc, b["is_real"])  # (the oracle stands in for qb200_residual)
    gathered = PAR.allgather_states(torch.from_numpy(res), nst).numpy()
    if rank == 0:
        c_all = R.synth_coefficients(b["kpg2"], ecut, nst, b["ngw"], b["is_real"], seed=3)
        h_all = R.synth_coefficients(b["kpg2"], ecut, nst, b["ngw"], b["is_real"], seed=4)
        res_all, _ = P.residual(c_all, h_all, b["is_real"])
        q.put((float(np.abs(call - c_all).max()), float(np.abs(gathered - res_all).max() / np.abs(res_all).max())))
    dist.barrier()
    dist.destroy_process_group()


def test_band_sharded_residual_world2():
    """a = c^H (Hc)_local needs every column of c: all-gather of the state blocks, then each rank's columns of the
    descent direction equal the single-rank result"""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 30100 + os.getpid() % 500
    procs = [ctx.Process(target=_worker_la, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0] == 0.0 and res[1] < 1e-13, res


def _worker_gram(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    here = os.path.dirname(os.path.abspath(__file__))
    for p in (os.path.dirname(here), os.path.join(os.path.dirname(here), "oracle"), here):
        if p not in sys.path:
            sys.path.insert(0, p)
    import port as P
    import refdrive as R
    dist.init_process_group("gloo", rank=rank, world_size=world)
    errs = []
    for kpoint, nst in (((0.25, 0, 0), 5), ((0, 0, 0), 7)):           # complex basis (blocks 3 + 2) and real basis (blocks 4 + 3)
        cell, ecut = (9, 0, 0, 0, 9, 0, 0, 0, 10), 4.0
        b = P.make_basis(cell, ecut, kpoint)
        first, n = PAR.state_block(nst, rank, world)
        c = R.synth_coefficients(b["kpg2"], ecut, n, b["ngw"], b["is_real"], seed=3, first_state=first)
        call = PAR.allgather_states(torch.from_numpy(c), nst).numpy()            # the gathered block qb200_gram_sharded takes
        # step 1 (qb200_gram_overlap): this rank's columns of S = c_all^H c_local, every other entry zero
        S = np.zeros((nst, nst), dtype=np.complex128)
        S[:, first:first + n] = P.subspace_h(call, c, b["is_real"])               # same product as the overlap: h with Hc := c
        # step 2 (qb200_allreduce_rho on the nall x nall block): every entry has one non-zero contributor
        St = torch.from_numpy(np.ascontiguousarray(S).view(np.float64).copy())
        dist.all_reduce(St)
        S = St.numpy().view(np.complex128).reshape(nst, nst)
        # step 3 (qb200_gram_apply): Cholesky replicated, own columns of c_all L^-H
        Lc = np.linalg.cholesky(np.tril(S) + np.tril(S, -1).conj().T)
        T = np.linalg.inv(Lc).conj().T                                             # L^-H, upper triangular
        mine = (T[:, first:first + n].T @ call)                                    # rows = this rank's new states
        want = P.gram(R.synth_coefficients(b["kpg2"], ecut, nst, b["ngw"], b["is_real"], seed=3), b["is_real"])[first:first + n]
        errs.append(float(np.abs(mine - want).max() / np.abs(want).max()) if n else 0.0)
    e = torch.tensor(errs, dtype=torch.float64)
    dist.all_reduce(e, op=dist.ReduceOp.MAX)
    if rank == 0:
        q.put(tuple(float(x) for x in e))
    dist.barrier()
    dist.destroy_process_group()


def test_band_sharded_gram_world2():
    """SlaterDet::gram over sharded states as qb200_gram_sharded runs it: overlap columns of the rank's block -> sum over the
    ranks -> Cholesky on every rank -> the rank's columns of c L^-H; the pieces reassemble the single-rank gram (numpy stands in
    for the device kernels; the collectives are the real ones)"""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 30400 + os.getpid() % 500
    procs = [ctx.Process(target=_worker_gram, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert max(res) < 1e-12, res
