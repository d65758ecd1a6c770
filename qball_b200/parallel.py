"""Band (state) parallelism across the GPUs of one box: the reference's process grid with nprow = 1, npcol = #GPUs
(Wavefunction.cc:262-271).  States are split in contiguous blocks of nb = ceil(nst/npcol) (SlaterDet.cc:228-231 with the
default blocking); v(r), projector tables and plan tables are replicated; H psi needs no exchange; the density needs ONE
all-reduce of rho(r) (ChargeDensity.cc:309, BLACS dsum over state columns) and the scalars (E_nl, NonLocalPotential.cc:2629;
integral of rho, ChargeDensity.cc:528) one small all-reduce.  torch.distributed (NCCL on GPUs, gloo in the CPU tests) is
the plumbing.
The subspace linear algebra (qb200_residual) is the one place with a real exchange step: every
rank needs ALL columns of c to form its columns of a = c^H (H c) and of c a -- allgather_states (NCCL all-gather over
NVLink) replaces the reference's pzgemm communication over process columns."""
from __future__ import annotations

import numpy as np  # noqa: F401  (host arrays are accepted by allreduce_density)
import torch
import torch.distributed as dist


def state_block(nst: int, rank: int, world: int):
    """(first, count) of the states owned by `rank`: block size nb = ceil(nst/world); trailing ranks may own fewer or none"""
    nb = nst // world + (1 if nst % world else 0)
    first = min(rank * nb, nst)
    return first, max(0, min(nb, nst - first))


def allreduce_density(rho, group=None):
    """wfcontext->dsum('r', np012loc, 1, rhor) over the state-column ranks (ChargeDensity.cc:309).  `rho` is a torch
    tensor (CUDA or CPU) or a numpy HOST array (the C ABI stages host pointers, so host grids are a supported input and
    must be summed over ranks exactly like device ones); reduced in place.  A host array under an NCCL-only process group
    is staged through the current CUDA device."""
    if not (dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1):
        return rho
    t = rho if isinstance(rho, torch.Tensor) else torch.from_numpy(rho)      # shares memory with the numpy array
    if not t.is_cuda and "gloo" not in str(dist.get_backend(group)).lower():
        if not torch.cuda.is_available():
            raise RuntimeError("allreduce_density: host array, but the process group has no CPU backend and no CUDA device")
        d = t.cuda()
        dist.all_reduce(d, op=dist.ReduceOp.SUM, group=group)
        t.copy_(d)
    else:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return rho


def allreduce_scalars(values, device=None, group=None):
    """sum a short list of python floats over ranks (E_nl partials, integral of rho, ...)"""
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return [float(x) for x in t.cpu()]


def allgather_states(c_local: torch.Tensor, nst: int, group=None) -> torch.Tensor:
    """(nst, ldc) block of ALL states from the per-rank blocks of state_block(); ranks that own fewer than nb states
    contribute zero columns that are trimmed again (the blocks are contiguous and in rank order)."""
    world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    if world == 1:
        return c_local
    nb = nst // world + (1 if nst % world else 0)
    ldc = c_local.shape[1]
    send = c_local
    if c_local.shape[0] != nb:
        send = torch.zeros((nb, ldc), dtype=c_local.dtype, device=c_local.device)
        send[:c_local.shape[0]] = c_local
    out = torch.empty((world * nb, ldc), dtype=c_local.dtype, device=c_local.device)
    # complex tensors travel as their (re, im) doubles
    dist.all_gather_into_tensor(torch.view_as_real(out), torch.view_as_real(send.contiguous()), group=group)
    return out[:nst]
