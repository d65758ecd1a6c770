// DMMA (mma.sync.m8n8k4.f64) issue-rate microbenchmark: W warps per SM, NACC independent accumulators per warp,
// optionally interleaved with plain FP64 FMAs (same pipe?).  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_peak dmma_peak.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b)
{
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
template <int NACC, int NFMA>
__global__ void k(double* out, int iters, double a0, double b0)
{
  double acc[NACC > 0 ? NACC : 1][2];
  double f[8];
  for (int i = 0; i < NACC; i++) acc[i][0] = acc[i][1] = 0.0;
  for (int i = 0; i < 8; i++) f[i] = threadIdx.x * 1e-3 + i;
  double a = a0 + threadIdx.x, b = b0 - threadIdx.x;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NACC; i++) dmma(acc[i][0], acc[i][1], a, b);
#pragma unroll
    for (int j = 0; j < NFMA; j++) f[j & 7] = fma(f[j & 7], a, b);
  }
  double s = 0;
  for (int i = 0; i < NACC; i++) s += acc[i][0] + acc[i][1];
  for (int i = 0; i < 8; i++) s += f[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int NACC, int NFMA> void run(int warps, const char* tag)
{
  double* out; cudaMalloc(&out, 148 * 1024 * 8);
  const int iters = 20000;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<NACC, NFMA><<<148, warps * 32>>>(out, 100, 1.0, 2.0);
  cudaEventRecord(e0);
  k<NACC, NFMA><<<148, warps * 32>>>(out, iters, 1.0, 2.0);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double flops = 2.0 * 256 * NACC * (double)iters * warps * 148;
  const double fmaf_ = 2.0 * 32 * NFMA * (double)iters * warps * 148;
  printf("%-28s warps/SM %2d  nacc %2d  nfma %2d : %8.3f ms  DMMA %6.2f TF/s  (+FMA %5.2f TF/s)  clk/DMMA/SMSP %.1f\n", tag, warps, NACC, NFMA, ms,
         flops / ms * 1e-9, fmaf_ / ms * 1e-9, ms * 1e-3 * 1.965e9 / ((double)NACC * iters * warps / 4.0));
  cudaFree(out);
}
int main()
{
  for (int w : {4, 8, 16, 32}) run<16, 0>(w, "dmma only");
  for (int w : {4, 8, 16}) run<4, 0>(w, "dmma only, 4 acc");
  for (int w : {8, 16}) run<16, 16>(w, "dmma + 16 fma / 16 dmma");
  for (int w : {8, 16}) run<16, 64>(w, "dmma + 64 fma / 16 dmma");
  for (int w : {8, 16}) run<0, 64>(w, "fma only");
  return 0;
}
