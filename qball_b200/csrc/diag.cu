// qball_b200/csrc/diag.cu -- device self-measurement used by bench.py for the FP64 roofline denominator.
// MEASURED_PEAKS.json (driver-written) carries HBM GB/s and bf16 TF/s only; the projector GEMMs run on the FP64 tensor
// path (mma.sync.m8n8k4.f64 = DMMA; tcgen05 has no FP64 kind), so their ceiling is measured here, in the same process:
// issue-rate loops of independent DMMA accumulators / independent DFMA chains on every SM.
#include "qb200_internal.h"

namespace qb200 {

__device__ __forceinline__ void dmma_acc(double& d0, double& d1, double a, double b)
{
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

template <int NACC, int NFMA>
__global__ void __launch_bounds__(512) k_fp64_rate(double* out, int iters, double a0, double b0)
{
  double acc[NACC > 0 ? NACC : 1][2];
  double f[NFMA > 0 ? NFMA : 1];
#pragma unroll
  for (int i = 0; i < NACC; i++) acc[i][0] = acc[i][1] = 0.0;
#pragma unroll
  for (int i = 0; i < NFMA; i++) f[i] = threadIdx.x * 1e-3 + i;
  const double a = a0 + 1e-9 * threadIdx.x, b = b0 - 1e-9 * threadIdx.x;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NACC; i++) dmma_acc(acc[i][0], acc[i][1], a, b);
#pragma unroll
    for (int j = 0; j < NFMA; j++) f[j] = fma(f[j], a, b);
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < NACC; i++) s += acc[i][0] + acc[i][1];
#pragma unroll
  for (int i = 0; i < NFMA; i++) s += f[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC, int NFMA>
static int rate(int nsm, int warps, int iters, double* out, double* tflops)
{
  cudaEvent_t e0, e1;
  QB_CUDA(cudaEventCreate(&e0));
  QB_CUDA(cudaEventCreate(&e1));
  k_fp64_rate<NACC, NFMA><<<nsm, warps * 32>>>(out, 64, 1.0, 2.0);        // warm-up
  double best = 0.0;
  for (int rep = 0; rep < 3; rep++) {
    QB_CUDA(cudaEventRecord(e0));
    k_fp64_rate<NACC, NFMA><<<nsm, warps * 32>>>(out, iters, 1.0, 2.0);
    QB_CUDA(cudaEventRecord(e1));
    QB_CUDA(cudaEventSynchronize(e1));
    float ms = 0.f;
    QB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    // one warp-wide DMMA m8n8k4 = 256 MACs; one warp-wide DFMA = 32 MACs
    const double fl = 2.0 * (256.0 * NACC + 32.0 * NFMA) * (double)iters * warps * nsm;
    best = fmax(best, fl / (ms * 1e-3) / 1e12);
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  *tflops = best;
  return QB200_OK;
}

}  // namespace qb200

using namespace qb200;

// out[0] = DMMA TFLOP/s, out[1] = DFMA TFLOP/s (best of 3 runs of ~10 ms each, 16 warps per SM, legacy default stream)
extern "C" int qb200_measure_fp64_peak(int device, double* out)
{
  if (!out) { set_error("qb200_measure_fp64_peak: bad argument"); return QB200_EINVAL; }
  int ndev = 0;
  QB_CUDA(cudaGetDeviceCount(&ndev));
  if (device < 0 || device >= ndev) { set_error("qb200_measure_fp64_peak: no such CUDA device"); return QB200_ENODEV; }
  QB_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  QB_CUDA(cudaGetDeviceProperties(&prop, device));
  const int nsm = prop.multiProcessorCount;
  double* buf = nullptr;
  QB_CUDA(cudaMalloc((void**)&buf, (size_t)nsm * 512 * sizeof(double)));
  int rc = rate<16, 0>(nsm, 16, 12000, buf, &out[0]);
  if (!rc) rc = rate<0, 32>(nsm, 16, 40000, buf, &out[1]);
  cudaFree(buf);
  return rc;
}
