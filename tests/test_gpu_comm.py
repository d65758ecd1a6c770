"""GPU: the C ABI's own collectives (qb200_comm_*, qb200_allreduce_rho / _scalars: NCCL inside libqball_b200.so) -- what a
C++ caller uses in place of ChargeDensity.cc:309's dsum('r') and the scalar dsums (NonLocalPotential.cc:2629,
EnergyFunctional.cc:1294, ChargeDensity.cc:528).  One process per GPU; the band-sharded density and E_nl reduced through
the library must equal the single-rank result over all states.  world_size 2 needs two GPUs (gpurun --gpus 2); with one
GPU the single-rank communicator still exercises NCCL loading, init and the no-op reductions."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    here = os.path.dirname(os.path.abspath(__file__))
    for p in (os.path.dirname(here), os.path.join(os.path.dirname(here), "oracle"), here):
        if p not in sys.path:
            sys.path.insert(0, p)
    import port as P
    import refdrive as R
    from qball_b200 import host as H
    from qball_b200 import parallel as PAR
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)          # rendezvous only: the id travels as an object
    comm = PAR.Communicator.from_torch_distributed(rank)
    assert comm.nranks == world and comm.nccl_version() >= 21800
    cell, ecut, nst = (9, 0, 0, 0, 9, 0, 0, 0, 10), 5.0, 7
    b = P.make_basis(cell, ecut, (0.25, 0, 0))
    grid = P.density_grid(cell, ecut)
    N = grid[0] * grid[1] * grid[2]
    occ = R.synth_occ(nst, 5)
    first, n = PAR.state_block(nst, rank, world)
    c = R.synth_coefficients(b["kpg2"], ecut, n, b["ngw"], b["is_real"], seed=3, first_state=first)
    rng = np.random.default_rng(0)
    sp = [dict(na=2, npr=4, lproj=np.array([0, 1, 1, 1], dtype=np.int32), wt=np.array([1.0, -0.5, -0.5, -0.5]),
               twnl=rng.standard_normal((4, b["ngw"])), tau=rng.uniform(0, 9, (2, 3)))]
    ft = H.FourierTransform(b, *grid, device=rank)
    nlp = H.NonLocalPotential(b, sp, device=rank)
    cd = torch.from_numpy(c).cuda()
    rho = torch.zeros(N, dtype=torch.float64, device="cuda")
    H.compute_density(ft, cd, 1.0, occ[first:first + n], b["omega"], rho)
    enl = nlp.energy(cd, occ[first:first + n], False)
    rho_host = rho.cpu().numpy().copy()
    nel_local = float(rho_host.sum()) * b["omega"] / N
    comm.allreduce_rho(rho)                                  # device pointer, on torch's current stream
    comm.allreduce_rho(rho_host)                             # host pointer, staged
    enl_sum, nel = comm.allreduce_scalars([enl, nel_local])
    torch.cuda.synchronize()
    # SlaterDet::gram over the sharded states: gathered block -> overlap columns -> sum over ranks -> replicated Cholesky -> own columns
    call_h = R.synth_coefficients(b["kpg2"], ecut, nst, b["ngw"], b["is_real"], seed=3)
    la = H.SubspaceLA(b, device=rank)
    c_all = torch.from_numpy(call_h).cuda()
    c_loc = torch.zeros_like(cd)
    la.gram_sharded(comm, c_all, first, n, c_loc)
    want = P.gram(call_h, b["is_real"])[first:first + n]
    gerr = torch.tensor([float(np.abs(c_loc.cpu().numpy() - want).max() / np.abs(want).max()) if n else 0.0], dtype=torch.float64)
    dist.all_reduce(gerr, op=dist.ReduceOp.MAX)
    la.close()
    if rank == 0:
        call = R.synth_coefficients(b["kpg2"], ecut, nst, b["ngw"], b["is_real"], seed=3)
        oft = P.FT(b, *grid)
        rho_all = oft.compute_density(call, occ / b["omega"], np.zeros(N))
        enl_all, _ = P.nl_energy(b, call, occ, sp, compute_hpsi=False)
        sc = np.abs(rho_all).max()
        q.put((float(np.abs(rho.cpu().numpy() - rho_all).max() / sc), float(np.abs(rho_host - rho_all).max() / sc),
               abs(enl_sum - enl_all) / max(1.0, abs(enl_all)), abs(nel - rho_all.sum() * b["omega"] / N), float(gerr[0])))
    dist.barrier()
    comm.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [1, 2])
def test_c_abi_collectives_band_sharded_density(world):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 30700 + os.getpid() % 500 + world
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = q.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert res[0] < 1e-10 and res[1] < 1e-10 and res[2] < 1e-10 and res[3] < 1e-9 and res[4] < 1e-10, res
