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


class Communicator:
    """qb200_comm: the NCCL communicator INSIDE libqball_b200.so (include/qball_b200.h, "collectives"), the object a C++
    caller gets.  The 128-byte unique id is produced by rank 0 and handed to the other ranks out of band -- here through
    torch.distributed's object broadcast (the reference would MPI_Bcast it over its own communicator)."""

    def __init__(self, device: int, rank: int, nranks: int, unique_id: bytes):
        import ctypes as C
        from . import capi
        self._L = capi.load()
        assert len(unique_id) == 128
        buf = C.create_string_buffer(unique_id, 128)
        h = C.c_void_p()
        capi._check(self._L.qb200_comm_init(C.byref(h), int(device), C.cast(buf, C.c_void_p), int(rank), int(nranks)), "qb200_comm_init")
        self._h, self.rank, self.nranks = h, rank, nranks

    @staticmethod
    def unique_id() -> bytes:
        import ctypes as C
        from . import capi
        buf = C.create_string_buffer(128)
        capi._check(capi.load().qb200_comm_get_unique_id(C.cast(buf, C.c_void_p)), "qb200_comm_get_unique_id")
        return buf.raw

    @classmethod
    def from_torch_distributed(cls, device: int, group=None):
        """one communicator over the ranks of an initialised torch.distributed group (rank 0's id broadcast as an object)"""
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        box = [cls.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0, group=group)
        return cls(device, rank, world, box[0])

    def nccl_version(self) -> int:
        return int(self._L.qb200_comm_query(self._h, 2))

    def allreduce_rho(self, rho, stream=None):
        """ChargeDensity.cc:309: sum of rho(r) over the state-column ranks, in place; device tensor (asynchronous on `stream`,
        default: torch's current stream) or host numpy array (staged, synchronous)"""
        from . import capi
        n = int(rho.numel()) if hasattr(rho, "numel") else int(rho.size)
        s = 0
        if hasattr(rho, "is_cuda") and rho.is_cuda:
            st = stream if stream is not None else torch.cuda.current_stream(rho.device)
            s = st.cuda_stream if hasattr(st, "cuda_stream") else int(st)
        capi._check(self._L.qb200_allreduce_rho(self._h, capi.ptr(rho), n, s), "qb200_allreduce_rho")
        return rho

    def allreduce_scalars(self, values):
        """NonLocalPotential.cc:2629 / EnergyFunctional.cc:1294 / ChargeDensity.cc:528: one small sum over the ranks"""
        from . import capi
        a = np.ascontiguousarray(np.array(list(values), dtype=np.float64))
        capi._check(self._L.qb200_allreduce_scalars(self._h, capi.ptr(a), int(a.size)), "qb200_allreduce_scalars")
        return [float(x) for x in a]

    def close(self):
        if getattr(self, "_h", None):
            self._L.qb200_comm_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
