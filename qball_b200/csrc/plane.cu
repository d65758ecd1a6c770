// qball_b200/csrc/plane.cu -- instantiation and launch of the plane-fused xy kernels (plane_kernels.cuh).
// Own translation unit: the __noinline__ radix passes are shared by all kernels of a translation unit and compiled
// for the tightest register budget among them; here that is 65536/448 = 144 registers (128 elsewhere).
#include "plane_kernels.cuh"

namespace qb200 {

int plane_opt_in(qb200_plan* p)
{
  const int bytes = (int)p->smem_plane;
  if (bytes <= 48 * 1024) return QB200_OK;
  QB_CUDA(cudaFuncSetAttribute(k_plane2<OP_HPSI>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  QB_CUDA(cudaFuncSetAttribute(k_plane2<OP_DENSITY>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  QB_CUDA(cudaFuncSetAttribute(k_plane2<OP_BWD>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  QB_CUDA(cudaFuncSetAttribute(k_plane2<OP_FWD>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  return QB200_OK;
}

int launch_plane(qb200_plan* p, int op, dim3 grid, const double* v, double* f, const double* fac, int nunits, int zero_imag)
{
  const DevPlan& d = p->d;
  cplx* zt = (cplx*)p->zt;
  switch (op) {
    case OP_HPSI: k_plane2<OP_HPSI><<<grid, p->plane_threads, p->smem_plane, p->stream>>>(d, zt, v, (cplx*)f, p->rho_part, fac, nunits, zero_imag); break;
    case OP_DENSITY: k_plane2<OP_DENSITY><<<grid, p->plane_threads, p->smem_plane, p->stream>>>(d, zt, v, (cplx*)f, p->rho_part, fac, nunits, zero_imag); break;
    case OP_BWD: k_plane2<OP_BWD><<<grid, p->plane_threads, p->smem_plane, p->stream>>>(d, zt, v, (cplx*)f, p->rho_part, fac, nunits, zero_imag); break;
    default: k_plane2<OP_FWD><<<grid, p->plane_threads, p->smem_plane, p->stream>>>(d, zt, v, (cplx*)f, p->rho_part, fac, nunits, zero_imag); break;
  }
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "k_plane2 launch", __FILE__, __LINE__);
  return QB200_OK;
}

}  // namespace qb200
