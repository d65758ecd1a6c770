// qball_b200/csrc/zcol_tmem.cu
// k_zcol_bwd_t / k_zcol_fwd_t: the sphere <-> column-form ends of the transform (vector_to_zvec + the z FFTs of bwd,
// FourierTransform.cc:1624-1665, 584-700; the z FFTs of fwd + 1/N + zvec_to_vector, :1300-1361, 1666-1683, fused with
// cp += ... + 0.5|k+G|^2 c, SlaterDet.cc:1027-1036, EnergyFunctional.cc:1675-1690) with ONE THREAD PER z-COLUMN and the
// column in the thread's own TENSOR-MEMORY lane (tcgen05.ld/st, 112 x 16 B = 448 of 512 columns), for complex bases on the
// compiled 112-plane shape (examples/MgO216).  Against k_zcol_bwd2/fwd2 (zcol_kernels.cuh):
//   * no column tile in shared memory: the transform passes move registers <-> TMEM, nothing is exchanged between threads,
//     twiddles are warp-uniform constants; shared memory only stages the block's coefficients (their order in the coefficient
//     block is rod by rod, one thread walks one rod);
//   * zt[unit][z][column] rows are written / read STRAIGHT from registers: a warp's 32 columns are 512 contiguous bytes;
//   * work items are (unit, block of 128 columns) dealt round-robin to one CTA per SM: no wave quantisation; the next item's
//     coefficients arrive by one TMA bulk copy (cp.async.bulk + mbarrier) while the current item is transformed.
// Index maps as in plane_tmem.cuh: z = 7a + b, k = k1 + 16 k2, TMEM slot (b, k1) at columns 4*(16 b + k1).
#include "qb200_internal.h"
#include "plane_static.cuh"
#include "tmem_ops.cuh"
#include "async_ops.cuh"
#include <algorithm>
#include <cmath>
#include <cstdlib>

namespace qb200 {

__constant__ double2 c_ztw[7 * 16];     // W_112^{b k1} (cos, sin)

#define ZT_COLS 128                      // columns per work item = TMEM lanes

template <int NP2_, int ZSPLIT_> struct ZTShape { static constexpr int NP2 = NP2_, ZSPLIT = ZSPLIT_, ZSKIP = NP2_ - 2 * ZSPLIT_; };
typedef ZTShape<112, 26> ZtMgO216;

struct ZItem { int unit, col0, ig0, cnt; };
__device__ __forceinline__ ZItem zitem(const DevPlan& P, int it, int nblk)
{
  ZItem r;
  r.unit = it / nblk;
  const int blk = it - r.unit * nblk;
  r.col0 = blk * ZT_COLS;
  const int c1 = min(r.col0 + ZT_COLS, P.nrods);
  r.ig0 = P.rod_first[r.col0];
  r.cnt = (c1 < P.nrods ? P.rod_first[c1] : P.ngw) - r.ig0;
  return r;
}

// ------------------------------------------------------------------------------------------------ backward
// grid (#SMs), block 128*MW (MW warps per TMEM lane quarter).  smem: stage[2][cmax] complex
template <class ZS, int MW>
__global__ void __launch_bounds__(128 * MW, 1) k_zcol_bwd_t(const __grid_constant__ DevPlan P, const cplx* __restrict__ c, size_t ldc,
                                                       cplx* __restrict__ zt, int nunits, int nblk, int cmax)
{
  static_assert(ZS::NP2 == 112, "thread-per-column passes are written for 112 = 16 x 7");
  constexpr int np2 = ZS::NP2;
  extern __shared__ __align__(128) unsigned char smraw[];
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(8) uint64_t mbar[2];
  cplx* stage = reinterpret_cast<cplx*>(smraw);
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
  if (warp == 0) tmem_alloc512(&tmem_slot);
  if (tid == 32) { mbar_init(&mbar[0], 1); mbar_init(&mbar[1], 1); }
  tmem_fence_before();
  __syncthreads();
  tmem_fence_after();
  const uint32_t tbase = tmem_slot;
  const int q = warp & 3, m = warp >> 2;
  const uint32_t t0 = tbase + ((uint32_t)(q * 32) << 16);
  const int blo = (7 * m) / MW, bhi = (7 * (m + 1)) / MW, klo = (16 * m) / MW, khi = (16 * (m + 1)) / MW;
  constexpr unsigned MASK = zmask(16, 7, ZS::ZSPLIT, ZS::ZSKIP);
  const int nitems = nunits * nblk;
  int it = blockIdx.x;
  auto issue = [&](const ZItem& w, int buf) {
    mbar_expect_tx(&mbar[buf], (uint32_t)w.cnt * 16u);
    bulk_g2s(stage + (size_t)buf * cmax, c + (size_t)w.unit * ldc + w.ig0, (uint32_t)w.cnt * 16u, &mbar[buf]);
  };
  if (it < nitems && tid == 0) issue(zitem(P, it, nblk), 0);
  uint32_t ph0 = 0, ph1 = 0;
  for (int buf = 0; it < nitems; it += gridDim.x, buf ^= 1) {
    const ZItem w = zitem(P, it, nblk);
    const int col = w.col0 + 32 * q + lane;
    const bool valid = col < P.nrods;
    int first = 0, size = 0, lmin = 0;
    if (valid) { first = P.rod_first[col] - w.ig0; size = P.rod_size[col]; lmin = P.rod_lmin[col]; }
    // everybody has left the previous item: its TMEM slots and the other staging buffer are free
    tmem_fence_before();
    __syncthreads();
    tmem_fence_after();
    if (tid == 0 && it + (int)gridDim.x < nitems) issue(zitem(P, it + gridDim.x, nblk), buf ^ 1);
    mbar_wait(&mbar[buf], buf ? ph1 : ph0);
    if (buf) ph1 ^= 1u; else ph0 ^= 1u;
    const cplx* rod = stage + (size_t)buf * cmax + first;
    // pass 1: 16-point transforms over a of the coefficients at z = 7a + b (zero outside the rod), twiddle -> slots (b, .)
#pragma unroll 1
    for (int b = blo; b < bhi; b++) {
      cplx x[16];
#pragma unroll
      for (int a = 0; a < 16; a++) {
        if (zclass(a, 7, ZS::ZSPLIT, ZS::ZSKIP) == 0) continue;
        const int zz = 7 * a + b;
        const int idx = (7 * a + 6 < np2 / 2 ? zz : (7 * a >= np2 / 2 ? zz - np2 : (zz < np2 / 2 ? zz : zz - np2))) - lmin;
        x[a] = ((unsigned)idx < (unsigned)size) ? rod[idx] : make_double2(0.0, 0.0);
      }
      DftM<16, +1, MASK>::run(x);
      if (b != 0) {
#pragma unroll
        for (int k1 = 1; k1 < 16; k1++) { const double2 tw = c_ztw[16 * b + k1]; x[k1] = cmul_s<+1>(x[k1], tw.x, tw.y); }
      }
      Tmem<16>::st(t0 + 64 * b, x);
    }
    tmem_wait_st();
    tmem_fence_before();
    bar_sync_n(1 + q, 32 * MW);
    tmem_fence_after();
    // pass 2: 7-point transforms over b -> the column at z = k1 + 16 k2, written to the z-major column form
    cplx* out = zt + (size_t)w.unit * np2 * P.nvec + min(col, P.nrods - 1);
#pragma unroll 1
    for (int k1 = klo; k1 < khi; k1++) {
      cplx t[7];
      Tmem<1, 7>::ld(t, t0 + 4 * k1, 64);
      Dft<7, +1>::run(t);
      if (valid) {
#pragma unroll
        for (int k2 = 0; k2 < 7; k2++) out[(size_t)(k1 + 16 * k2) * P.nvec] = t[k2];
      }
    }
  }
  tmem_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc512(tbase);
}

// ------------------------------------------------------------------------------------------------ forward
// grid (#SMs), block 512 (four warps per TMEM lane quarter).  smem: stage[cmax] complex
template <class ZS>
__global__ void __launch_bounds__(512, 1) k_zcol_fwd_t(const __grid_constant__ DevPlan P, const cplx* __restrict__ zt, cplx* __restrict__ out,
                                                       size_t ldc, int accumulate, const double* __restrict__ kpg2,
                                                       const cplx* __restrict__ cin, double scale, int nunits, int nblk, int cmax)
{
  static_assert(ZS::NP2 == 112, "thread-per-column passes are written for 112 = 16 x 7");
  constexpr int np2 = ZS::NP2;
  extern __shared__ __align__(128) unsigned char smraw[];
  __shared__ uint32_t tmem_slot;
  cplx* stage = reinterpret_cast<cplx*>(smraw);
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
  if (warp == 0) tmem_alloc512(&tmem_slot);
  tmem_fence_before();
  __syncthreads();
  tmem_fence_after();
  const uint32_t tbase = tmem_slot;
  const int q = warp & 3, m = warp >> 2;
  const uint32_t t0 = tbase + ((uint32_t)(q * 32) << 16);
  constexpr int MW = 4;
  const int blo = (7 * m) / MW, bhi = (7 * (m + 1)) / MW, klo = (16 * m) / MW;
  const int nitems = nunits * nblk;
  for (int it = blockIdx.x; it < nitems; it += gridDim.x) {
    const ZItem w = zitem(P, it, nblk);
    const int col = w.col0 + 32 * q + lane;
    const bool valid = col < P.nrods;
    int first = 0, size = 0, lmin = 0;
    if (valid) { first = P.rod_first[col] - w.ig0; size = P.rod_size[col]; lmin = P.rod_lmin[col]; }
    // everybody has left the previous item: its TMEM slots and the staging buffer are free
    tmem_fence_before();
    __syncthreads();
    tmem_fence_after();
    // pass A: 7-point transforms over k2 of the column values at z = k1 + 16 k2 (read straight from zt), twiddle -> slots (., k1)
    const cplx* in = zt + (size_t)w.unit * np2 * P.nvec + min(col, P.nrods - 1);
#pragma unroll 1
    for (int kk = 0; kk < 4; kk += 2) {
      cplx t[2][7];
#pragma unroll
      for (int j = 0; j < 2; j++)
#pragma unroll
        for (int k2 = 0; k2 < 7; k2++) t[j][k2] = __ldcs(in + (size_t)(klo + kk + j + 16 * k2) * P.nvec);
#pragma unroll
      for (int j = 0; j < 2; j++) {
        const int k1 = klo + kk + j;
        Dft<7, -1>::run(t[j]);
#pragma unroll
        for (int b = 1; b < 7; b++) { const double2 tw = c_ztw[16 * b + k1]; t[j][b] = cmul_s<-1>(t[j][b], tw.x, tw.y); }
        Tmem<1, 7>::st(t0 + 4 * k1, t[j], 64);
      }
    }
    tmem_wait_st();
    tmem_fence_before();
    bar_sync_n(1 + q, 32 * MW);
    tmem_fence_after();
    // pass B: 16-point transforms over k1 -> z = 7a + b; the rod's coefficients go to their place in the staged block
    cplx* rod = stage + first;
#pragma unroll 1
    for (int b = blo; b < bhi; b++) {
      cplx x[16];
      Tmem<16>::ld(x, t0 + 64 * b);
      Dft<16, -1>::run(x);
#pragma unroll
      for (int a = 0; a < 16; a++) {
        if (zclass(a, 7, ZS::ZSPLIT, ZS::ZSKIP) == 0) continue;
        const int zz = 7 * a + b;
        const int idx = (7 * a + 6 < np2 / 2 ? zz : (7 * a >= np2 / 2 ? zz - np2 : (zz < np2 / 2 ? zz : zz - np2))) - lmin;
        if ((unsigned)idx < (unsigned)size) rod[idx] = x[a];
      }
    }
    __syncthreads();
    // epilogue over the block's coefficients, coalesced: cp (+)= scale * f + 0.5 |k+G|^2 c
    const size_t s1 = (size_t)w.unit * ldc + w.ig0;
    cplx* o1 = out + s1;
    const cplx* i1 = cin + s1;
    const double* kp = kpg2 + w.ig0;
    constexpr int U = 4;
    for (int e0 = tid; e0 < w.cnt; e0 += U * 512) {
      cplx a1[U], q1[U];
      double kh[U];
#pragma unroll
      for (int u = 0; u < U; u++) {
        const int e = e0 + u * 512;
        if (e < w.cnt) {
          if (kpg2) { a1[u] = i1[e]; kh[u] = 0.5 * kp[e]; }
          if (accumulate) q1[u] = o1[e];
        }
      }
#pragma unroll
      for (int u = 0; u < U; u++) {
        const int e = e0 + u * 512;
        if (e < w.cnt) {
          const cplx pv = stage[e];
          cplx w1 = make_double2(scale * pv.x, scale * pv.y);
          if (kpg2) { w1.x += kh[u] * a1[u].x; w1.y += kh[u] * a1[u].y; }
          if (accumulate) { w1.x += q1[u].x; w1.y += q1[u].y; }
          o1[e] = w1;
        }
      }
    }
  }
  tmem_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc512(tbase);
}

// ------------------------------------------------------------------------------------------------ host side
bool zcol_t_wanted(const qb200_plan* p, int lmax)
{
  if (const char* e = getenv("QB200_ZCOL_T")) if (e[0] == '0') return false;
  if (const char* e = getenv("QB200_NO_STATIC")) if (e[0] == '1') return false;
  const DevPlan& d = p->d;
  return !d.is_real && d.np2 == ZtMgO216::NP2 && lmax < ZtMgO216::ZSPLIT;
}

int zcol_t_setup(qb200_plan* p, const std::vector<int>& first)
{
  const DevPlan& d = p->d;
  p->zcol_t = false;
  int cmax = 4;
  for (int r0 = 0; r0 < d.nrods; r0 += ZT_COLS) {
    const int r1 = std::min(r0 + ZT_COLS, d.nrods);
    cmax = std::max(cmax, (r1 < d.nrods ? first[r1] : d.ngw) - first[r0]);
  }
  cmax = (cmax + 7) & ~7;
  p->zt_cmax = cmax;
  p->zt_nblk = (d.nrods + ZT_COLS - 1) / ZT_COLS;
  p->smem_zt_b = 2 * (size_t)cmax * 16;
  p->smem_zt_f = (size_t)cmax * 16;
  if (p->smem_zt_b + 256 > (size_t)p->max_smem) return QB200_OK;       // blocks too long to double-buffer: the v2 kernels stay
  double tw[2 * 7 * 16];
  const long double twopi = 6.283185307179586476925286766559005768L;
  for (int b = 0; b < 7; b++)
    for (int k1 = 0; k1 < 16; k1++) {
      const int e = (b * k1) % ZtMgO216::NP2;
      tw[2 * (16 * b + k1)] = (double)cosl(twopi * e / ZtMgO216::NP2);
      tw[2 * (16 * b + k1) + 1] = (double)sinl(twopi * e / ZtMgO216::NP2);
    }
  QB_CUDA(cudaMemcpyToSymbol(c_ztw, tw, sizeof(tw)));
  QB_CUDA((cudaFuncSetAttribute(k_zcol_bwd_t<ZtMgO216, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p->smem_zt_b)));
  QB_CUDA((cudaFuncSetAttribute(k_zcol_bwd_t<ZtMgO216, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p->smem_zt_b)));
  QB_CUDA(cudaFuncSetAttribute(k_zcol_fwd_t<ZtMgO216>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p->smem_zt_f));
  p->zcol_t = true;
  return QB200_OK;
}

int launch_zbwd_t(qb200_plan* p, const double* c, size_t ldc, int nunits)
{
  const int nitems = nunits * p->zt_nblk;
  static const int mw = [] { const char* e = getenv("QB200_ZB_MW"); return e ? atoi(e) : 2; }();
  if (mw == 4) k_zcol_bwd_t<ZtMgO216, 4><<<std::min(p->nsm, nitems), 512, p->smem_zt_b, p->stream>>>(p->d, (const cplx*)c, ldc, (cplx*)p->zt, nunits, p->zt_nblk, p->zt_cmax);
  else k_zcol_bwd_t<ZtMgO216, 2><<<std::min(p->nsm, nitems), 256, p->smem_zt_b, p->stream>>>(p->d, (const cplx*)c, ldc, (cplx*)p->zt, nunits, p->zt_nblk, p->zt_cmax);
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "k_zcol_bwd_t launch", __FILE__, __LINE__);
  return QB200_OK;
}

int launch_zfwd_t(qb200_plan* p, double* out, size_t ldc, int nunits, int accumulate, const double* kpg2, const double* cin, double scale)
{
  const int nitems = nunits * p->zt_nblk;
  k_zcol_fwd_t<ZtMgO216><<<std::min(p->nsm, nitems), 512, p->smem_zt_f, p->stream>>>(p->d, (const cplx*)p->zt, (cplx*)out, ldc, accumulate, kpg2,
                                                                                  (const cplx*)cin, scale, nunits, p->zt_nblk, p->zt_cmax);
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "k_zcol_fwd_t launch", __FILE__, __LINE__);
  return QB200_OK;
}

}  // namespace qb200
