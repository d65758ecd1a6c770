// tools/microbench/tmem_ldst.cu -- tensor memory (TMEM) as a lane-private workspace for FP64 butterflies: does it work, what
// does tcgen05.ld / tcgen05.st cost on B200, and do they overlap with the FP64 pipe and with shared-memory traffic?
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/microbench/tmem_ldst tools/microbench/tmem_ldst.cu
// Output: one JSON object (bytes per cycle per SM for each pattern, correctness flags).
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../qball_b200/csrc/tmem_ops.cuh"
using namespace qb200;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("{\"error\": \"%s at %s:%d\"}\n", cudaGetErrorString(e), __FILE__, __LINE__); return 1; } } while (0)

__device__ __forceinline__ uint32_t lane_base(uint32_t base) { return base + ((uint32_t)((threadIdx.x >> 5) & 3) * 32u << 16); }

// mode 0: st.x64 stream; 1: ld.x64 stream; 2: ld 7 x (x8) strided by 64 columns; 3: st 7 x (x8); 4: ld x64 + st x64 alternating
// 5: FP64 only (256 DFMA per iteration); 6: ld x64 + 256 DFMA + st x64; 7: as 6 for even warps, odd warps stream LDS/STS 16-byte
// 8: odd warps LDS/STS only (even idle)
template <int MODE>
__global__ void __launch_bounds__(512, 1) k_bench(int iters, double* out, long long* cycles)
{
  __shared__ uint32_t slot;
  extern __shared__ __align__(16) unsigned char smraw[];
  double2* sm = reinterpret_cast<double2*>(smraw);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) tmem_alloc512(&slot);
  tmem_fence_before();
  __syncthreads();
  tmem_fence_after();
  const uint32_t base = slot;
  // warps sharing a lane quarter use disjoint column ranges when there are more than 4 warps
  const int nshare = (blockDim.x / 32 + 3) / 4, share = warp / 4;
  const uint32_t colspan = 512 / nshare;           // columns this warp may touch
  const uint32_t t0 = lane_base(base) + share * colspan;
  double2 x[16];
#pragma unroll
  for (int i = 0; i < 16; i++) x[i] = make_double2(1.0 + lane + i, 0.5 * warp + i);
  // initialise the warp's range so that loads read defined data
  for (uint32_t c = 0; c + 64 <= colspan; c += 64) Tmem<16>::st(t0 + c, x);
  tmem_wait_st();
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = make_double2(i, -i);
  __syncthreads();
  const long long c0 = clock64();
  const uint32_t nblk = colspan / 64;
  double acc = 0.0;
  if (MODE == 0) {
    for (int it = 0; it < iters; it++) Tmem<16>::st(t0 + (it % nblk) * 64, x);
    tmem_wait_st();
  } else if (MODE == 1) {
    for (int it = 0; it < iters; it++) { double2 y[16]; Tmem<16>::ld(y, t0 + (it % nblk) * 64); acc += y[0].x + y[15].y; }
  } else if (MODE == 2) {
    if (colspan >= 448)
      for (int it = 0; it < iters; it++) { double2 y[14]; Tmem<2, 7>::ld(y, t0 + (it & 7) * 8, 64); acc += y[0].x + y[13].y; }
  } else if (MODE == 3) {
    if (colspan >= 448) {
      for (int it = 0; it < iters; it++) Tmem<2, 7>::st(t0 + (it & 7) * 8, x, 64);
      tmem_wait_st();
    }
  } else if (MODE == 4) {
    for (int it = 0; it < iters; it++) {
      double2 y[16];
      Tmem<16>::ld(y, t0 + (it % nblk) * 64);
#pragma unroll
      for (int i = 0; i < 16; i++) { y[i].x += 1.0; }
      Tmem<16>::st(t0 + (it % nblk) * 64, y);
      tmem_wait_st();
    }
  } else if (MODE == 5 || MODE == 6 || MODE == 7 || MODE == 8) {
    const bool fp = (MODE == 5 || MODE == 6) || (MODE == 7 && (warp & 1) == 0);
    const bool tm = (MODE == 6) || (MODE == 7 && (warp & 1) == 0);
    const bool ls = (MODE == 7 || MODE == 8) && (warp & 1) == 1;
    if (fp) {
      for (int it = 0; it < iters; it++) {
        double2 y[16];
        if (tm) Tmem<16>::ld(y, t0 + (it % nblk) * 64);
        else {
#pragma unroll
          for (int i = 0; i < 16; i++) y[i] = x[i];
        }
#pragma unroll
        for (int r = 0; r < 8; r++)
#pragma unroll
          for (int i = 0; i < 16; i++) { y[i].x = fma(y[i].x, 1.0000001, y[(i + 1) & 15].y); y[i].y = fma(y[i].y, 0.9999999, y[(i + 5) & 15].x); }
        if (tm) { Tmem<16>::st(t0 + (it % nblk) * 64, y); tmem_wait_st(); }
        else {
#pragma unroll
          for (int i = 0; i < 16; i++) x[i] = y[i];
        }
      }
#pragma unroll
      for (int i = 0; i < 16; i++) acc += x[i].x + x[i].y;
    }
    if (ls) {
      double2* p = sm + (warp >> 1) * 512 + lane;
      for (int it = 0; it < iters; it++) {
        double2 y[16];
#pragma unroll
        for (int i = 0; i < 16; i++) y[i] = p[i * 32];
#pragma unroll
        for (int i = 0; i < 16; i++) p[i * 32] = make_double2(y[i].y, y[i].x);
      }
      acc += p[0].x;
    }
  }
  const long long c1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) cycles[blockIdx.x] = c1 - c0;
  if (acc == 1.2345e301) out[0] = acc;
  tmem_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc512(base);
}

// correctness: (a) a thread reads back what it stored (all Tmem<> shapes, dynamic addresses); (b) warp w+4 reads what warp w
// stored after fence / barrier / fence (two warps of one lane quarter cooperating on one column of data)
__global__ void __launch_bounds__(256, 1) k_check(int* bad)
{
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) tmem_alloc512(&slot);
  tmem_fence_before();
  __syncthreads();
  tmem_fence_after();
  const uint32_t base = slot;
  const uint32_t t0 = lane_base(base);
  int nbad = 0;
  const int q = warp & 3;
  auto val = [&](int s) { return make_double2(1000.0 * (q * 32 + lane) + s, -0.25 * s + lane); };
  if (warp < 4) {                      // slots 0..111 of every lane (one column of a 112-point line), written 16 at a time
    for (int b = 0; b < 7; b++) {
      double2 x[16];
      for (int i = 0; i < 16; i++) x[i] = val(16 * b + i);
      Tmem<16>::st(t0 + 64 * b, x);
    }
    tmem_wait_st();
  }
  tmem_fence_before();
  __syncthreads();
  tmem_fence_after();
  if (warp >= 4) {                     // the other warp of the quarter reads with the strided pattern, modifies, writes back
    for (int k1 = 0; k1 < 16; k1 += 2) {
      double2 y[14];
      Tmem<2, 7>::ld(y, t0 + 4 * k1, 64);
      for (int b = 0; b < 7; b++)
        for (int j = 0; j < 2; j++) {
          const double2 e = val(16 * b + k1 + j);
          if (y[2 * b + j].x != e.x || y[2 * b + j].y != e.y) nbad++;
          y[2 * b + j].x += 7.0;
        }
      Tmem<2, 7>::st(t0 + 4 * k1, y, 64);
    }
    tmem_wait_st();
  }
  tmem_fence_before();
  __syncthreads();
  tmem_fence_after();
  if (warp < 4) {
    for (int b = 0; b < 7; b++) {
      double2 x[16];
      Tmem<16>::ld(x, t0 + 64 * b);
      for (int i = 0; i < 16; i++) { const double2 e = val(16 * b + i); if (x[i].x != e.x + 7.0 || x[i].y != e.y) nbad++; }
    }
    for (int s = 0; s < 112; s += 8) {   // other widths
      double2 x[8];
      Tmem<8>::ld(x, t0 + 4 * s);
      for (int i = 0; i < 8; i++) { const double2 e = val(s + i); if (x[i].x != e.x + 7.0 || x[i].y != e.y) nbad++; }
      double2 y4[4];
      Tmem<4>::ld(y4, t0 + 4 * s + 16);
      for (int i = 0; i < 4; i++) { const double2 e = val(s + 4 + i); if (y4[i].x != e.x + 7.0 || y4[i].y != e.y) nbad++; }
      double2 y1[1];
      Tmem<1>::ld(y1, t0 + 4 * s + 12);
      { const double2 e = val(s + 3); if (y1[0].x != e.x + 7.0 || y1[0].y != e.y) nbad++; }
    }
  }
  if (nbad) atomicAdd(bad, nbad);
  tmem_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc512(base);
}

template <int MODE> static double run(int nthreads, int iters, double bytes_per_iter_per_warp, int active_warps_div, const char* name, bool last = false)
{
  int nsm = 148;
  cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0); nsm = prop.multiProcessorCount;
  long long* cyc; double* out;
  cudaMalloc(&cyc, nsm * sizeof(long long)); cudaMalloc(&out, 8);
  cudaFuncSetAttribute(k_bench<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  k_bench<MODE><<<nsm, nthreads, 65536>>>(10, out, cyc);
  cudaDeviceSynchronize();
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k_bench<MODE><<<nsm, nthreads, 65536>>>(iters, out, cyc);
  cudaEventRecord(e1);
  cudaError_t e = cudaDeviceSynchronize();
  float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
  std::vector<long long> h(nsm);
  cudaMemcpy(h.data(), cyc, nsm * sizeof(long long), cudaMemcpyDeviceToHost);
  double avg = 0; for (long long c : h) avg += (double)c; avg /= nsm;
  const int nw = nthreads / 32 / active_warps_div;
  const double bpc = bytes_per_iter_per_warp * iters * nw / avg;
  printf(" \"%s_w%d\": {\"cycles_per_iter\": %.1f, \"bytes_per_cycle_per_sm\": %.1f, \"ms\": %.3f, \"err\": \"%s\"}%s\n", name, nthreads / 32, avg / iters, bpc, ms,
         e == cudaSuccess ? "" : cudaGetErrorString(e), last ? "" : ",");
  cudaFree(cyc); cudaFree(out);
  return bpc;
}

int main()
{
  int* bad; CK(cudaMalloc(&bad, 4)); CK(cudaMemset(bad, 0, 4));
  k_check<<<148, 256>>>(bad);
  CK(cudaDeviceSynchronize());
  int hb = -1; CK(cudaMemcpy(&hb, bad, 4, cudaMemcpyDeviceToHost));
  printf("{\n \"check_mismatches\": %d,\n", hb);
  const int it = 20000;
  const double B = 32 * 256.0;     // bytes per warp per x64 access
  for (int nt : { 128, 256, 512 }) {
    run<0>(nt, it, B, 1, "st_x64");
    run<1>(nt, it, B, 1, "ld_x64");
    run<4>(nt, it, 2 * B, 1, "ld_st_x64");
  }
  run<2>(128, it, 32 * 224.0, 1, "ld_7x8");
  run<3>(128, it, 32 * 224.0, 1, "st_7x8");
  for (int nt : { 256, 512 }) {
    run<5>(nt, it / 4, 256 * 32 * 8.0, 1, "dfma256_only(bytes=flops/2)");
    run<6>(nt, it / 4, 2 * B, 1, "ld_dfma256_st");
    run<7>(nt, it / 4, 2 * B, 2, "even:ld_dfma256_st_odd:lds_sts");
    run<8>(nt, it / 4, 32 * 512.0, 2, "odd:lds_sts_only(smem bytes)");
  }
  printf(" \"done\": 1\n}\n");
  return 0;
}
