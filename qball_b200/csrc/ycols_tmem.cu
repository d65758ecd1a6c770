// qball_b200/csrc/ycols_tmem.cu
// k_ycols_t<OP, YS>: the y stage of the SPLIT xy path (planes that do not fit one SM's shared memory: Au992 252 x 252, si54p
// 126 x 126) with the column in TENSOR MEMORY -- the round trip  y-transform(+1) -> v(r) multiply or |psi|^2 -> y-transform(-1)
// of FourierTransform.cc:822-974 / 1156-1298 around SlaterDet.cc:919-921, 993-1031 -- replacing k_ycols2 (split_kernels.cuh),
// whose three radix passes per direction ran through shared memory (ncu: shared-memory pipe 66 % busy, DRAM 830 GB/s).
// Here the kept rows w[unit][z][jr][x] go straight from HBM into registers (lanes = consecutive x: coalesced), every pass moves
// registers <-> the thread's own TMEM lane (tcgen05.st/ld), and the result returns to w straight from registers: no shared
// memory on the data path at all.
//
// A TMEM lane holds 128 complex doubles, so the per-lane transform has length NL = 126 = 9 x 14:
//   y' = 14a + b, k = k1 + 9 k2:   X[k1 + 9 k2] = sum_b W_14^{S b k2} [ W_126^{S b k1} sum_a W_9^{S a k1} x[14a + b] ]
//   (slot (b, k1) = 32-bit columns 4*(9 b + k1)), and transposed for the way back.
//   * np1 = 126 (si54p): one lane per column, 126 columns = one plane per work item.
//   * np1 = 252 (Au992): TWO lanes per column (lane and lane + 16 of a warp).  First radix-2 step by hand: with
//     u[y'] = x[y'] + x[y'+126], v[y'] = (x[y'] - x[y'+126]) W_252^{y'} the even outputs are the 126-point transform of u,
//     the odd ones that of v; only rows y < 56 or y >= 196 are non-zero, so u[y'] and v[y'] are ONE input each (no add):
//     both lanes read the same kept rows, the "odd" lane multiplies by +-W_252^{y'}.  On the way back
//     Y[y'] = u'[y'] + W^{-y'} v'[y'] (rows y < 126), Y[y'+126] = u'[y'] - W^{-y'} v'[y'] is one lane exchange (shuffle).
// Every (plane z, block of columns) has ONE owner CTA (persistent, one per SM) that walks the units in order.  The density is
// accumulated in shared memory over all units of a launch (acc[y][column], 127 KB: each point has one owner thread) and added to
// rho once per launch -- one read-modify-write per point and batch instead of one L2 reduction per point and state; fixed
// order, deterministic, as in the other density kernels.
#include "qb200_internal.h"
#include "plane_static.cuh"
#include "tmem_ops.cuh"
#include "async_ops.cuh"
#include <algorithm>
#include <cmath>
#include <cstdlib>

namespace qb200 {

__constant__ double2 c_w126[14 * 9];     // W_126^{b k1} at [9 b + k1]
__constant__ double2 c_w252[126];        // W_252^{y'}

// YS: NP1 = plane height (126 or 252); KSPLIT/KSKIP describe the non-zero per-lane inputs y' in [0,KSPLIT) and [KSPLIT+KSKIP,126)
template <int NP0_, int NP1_, int YSPLIT_> struct YTShape {
  static constexpr int NP0 = NP0_, NP1 = NP1_, YSPLIT = YSPLIT_, NKEEP = 2 * YSPLIT_;
  static constexpr bool PAIR = NP1_ == 252;
  static constexpr int NL = 126;
  static constexpr int KSPLIT = YSPLIT_, KSKIP = PAIR ? (NP1_ - 2 * YSPLIT_) - 126 : NP1_ - 2 * YSPLIT_;   // 252: [56,70) empty; 126: [29,97)
  static constexpr int COLS = PAIR ? 63 : 126;                      // columns per work item (126 active TMEM lanes)
  static constexpr int NXB = (NP0_ + COLS - 1) / COLS;
  static_assert(NP1_ == 126 || NP1_ == 252, "per-lane transform length is 126");
  static_assert(KSKIP >= 0 && 2 * YSPLIT_ <= NP1_, "kept rows");
};
typedef YTShape<126, 126, 29> YtSi54p;      // examples/si54p at 65 Ry
typedef YTShape<252, 252, 56> YtAu992;      // examples/gold_benchmark

template <int OP, class YS>
__global__ void __launch_bounds__(512, 1) k_ycols_t(const __grid_constant__ DevPlan P, cplx* __restrict__ w, const double* __restrict__ v,
                                                    double* __restrict__ rho_part, const double* __restrict__ fac, int nunits, int zero_imag)
{
  static_assert(OP == OP_HPSI || OP == OP_DENSITY, "k_ycols_t: H psi and density only");
  constexpr int np0 = YS::NP0, np1 = YS::NP1, np01 = np0 * np1, NK = YS::NKEEP;
  constexpr bool PAIR = YS::PAIR;
  constexpr unsigned MASK = zmask(9, 14, YS::KSPLIT, YS::KSKIP);
  constexpr int MW = 4;                                             // warps per TMEM lane quarter
  __shared__ uint32_t tmem_slot;
  extern __shared__ __align__(16) unsigned char smraw[];
  double* acc = reinterpret_cast<double*>(smraw);                  // DENSITY: [np1][COLS]
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
  if (warp == 0) tmem_alloc512(&tmem_slot);
  tmem_fence_before();
  __syncthreads();
  tmem_fence_after();
  const uint32_t tbase = tmem_slot;
  const int q = warp & 3, m = warp >> 2;
  const uint32_t t0 = tbase + ((uint32_t)(q * 32) << 16);
  const int blo = (14 * m) / MW, bhi = (14 * (m + 1)) / MW, klo = (9 * m) / MW, khi = (9 * (m + 1)) / MW;
  // column of this lane inside a work item, and its role (0: even outputs / plain column, 1: odd outputs)
  const int cl = PAIR ? 16 * q + (lane & 15) : 32 * q + lane;
  const int role = PAIR ? (lane >> 4) : 0;
  const int nslots = P.np2 * YS::NXB;                               // (z, block) pairs: each has one owner CTA
  for (int slot = blockIdx.x; slot < nslots; slot += gridDim.x) {
    const int z = slot / YS::NXB, xb = slot - z * YS::NXB;
    const int x = xb * YS::COLS + cl;
    const bool act = cl < YS::COLS && x < np0;
    const int xc = min(x, np0 - 1);
    const double* vz = v + (size_t)z * np01 + xc;
    if (OP == OP_DENSITY) {
      __syncthreads();                                             // the previous slot's write-out is done with acc
      for (int i = tid; i < np1 * YS::COLS; i += 512) acc[i] = 0.0;
      __syncthreads();
    }
    for (int unit = 0; unit < nunits; unit++) {
      double facu = 0.0, facv = 0.0;
      if (OP == OP_DENSITY) { if (!fac_active(P, fac, unit)) continue; facu = fac_first(fac, unit); facv = fac_second(P, fac, unit); }
      cplx* wz = w + ((size_t)unit * P.np2 + z) * NK * np0 + xc;
      // the previous item's pass 3 (other warps of the quarter) is done with the TMEM slots
      tmem_fence_before();
      bar_sync_n(1 + q, 32 * MW);
      tmem_fence_after();
      // pass 1: 9-point transforms over a of the kept rows y' = 14a + b, twiddle -> slots (b, .)
#pragma unroll 1
      for (int b = blo; b < bhi; b++) {
        cplx xin[9];
#pragma unroll
        for (int a = 0; a < 9; a++) {
          const int c = zclass(a, 14, YS::KSPLIT, YS::KSKIP);
          if (c == 0) continue;
          const int yp = 14 * a + b;
          const bool kept = c == 1 || yp < YS::KSPLIT || yp >= YS::KSPLIT + YS::KSKIP;
          const int jr = (14 * a + 13 < YS::KSPLIT) ? yp : ((14 * a >= YS::KSPLIT + YS::KSKIP) ? yp - YS::KSKIP : (yp < YS::KSPLIT ? yp : yp - YS::KSKIP));
          cplx val = make_double2(0.0, 0.0);
          if (kept) val = wz[(size_t)jr * np0];
          if (PAIR) {
            // odd lane: v[y'] = +x[y'] W^{y'} below the gap, -x[y'+126] W^{y'} above it
            const double2 tw = c_w252[yp];
            const double sg = (yp < YS::KSPLIT) ? 1.0 : -1.0;
            const cplx o = cmul_s<+1>(val, sg * tw.x, sg * tw.y);
            val = role ? o : val;
          }
          xin[a] = val;
        }
        DftM<9, +1, MASK>::run(xin);
        if (b != 0) {
#pragma unroll
          for (int k1 = 1; k1 < 9; k1++) { const double2 tw = c_w126[9 * b + k1]; xin[k1] = cmul_s<+1>(xin[k1], tw.x, tw.y); }
        }
        Tmem<8>::st(t0 + 36 * b, xin);
        Tmem<1>::st(t0 + 36 * b + 32, xin + 8);
      }
      tmem_wait_st();
      tmem_fence_before();
      bar_sync_n(1 + q, 32 * MW);
      tmem_fence_after();
      // pass 2: 14-point transforms over b -> psi(x, y, z) at y = k (one lane per column) or y = 2k + role; pointwise work; way back
#pragma unroll 1
      for (int k1 = klo; k1 < khi; k1++) {
        cplx t[14];
        Tmem<1, 14>::ld(t, t0 + 4 * k1, 36);
        Dft<14, +1>::run(t);
        if (OP == OP_HPSI) {
#pragma unroll
          for (int k2 = 0; k2 < 14; k2++) {
            const int k = k1 + 9 * k2, y = PAIR ? 2 * k + role : k;
            const double vv = __ldg(vz + (size_t)y * np0);
            t[k2].x *= vv;
            t[k2].y = zero_imag ? 0.0 : t[k2].y * vv;
          }
          Dft<14, -1>::run(t);
#pragma unroll
          for (int b = 1; b < 14; b++) { const double2 tw = c_w126[9 * b + k1]; t[b] = cmul_s<-1>(t[b], tw.x, tw.y); }
          Tmem<1, 14>::st(t0 + 4 * k1, t, 36);
        } else if (act) {
#pragma unroll
          for (int k2 = 0; k2 < 14; k2++) {
            const int k = k1 + 9 * k2, y = PAIR ? 2 * k + role : k;
            acc[y * YS::COLS + cl] += facu * t[k2].x * t[k2].x + facv * t[k2].y * t[k2].y;    // this thread owns (y, cl) in every unit
          }
        }
      }
      if (OP == OP_HPSI) {
        tmem_wait_st();
        tmem_fence_before();
        bar_sync_n(1 + q, 32 * MW);
        tmem_fence_after();
        // pass 3: 9-point transforms over k1 -> the kept rows y' = 14a + b
#pragma unroll 1
        for (int b = blo; b < bhi; b++) {
          cplx xo[9];
          Tmem<8>::ld(xo, t0 + 36 * b);
          Tmem<1>::ld(xo + 8, t0 + 36 * b + 32);
          Dft<9, -1>::run(xo);
#pragma unroll
          for (int a = 0; a < 9; a++) {
            const int c = zclass(a, 14, YS::KSPLIT, YS::KSKIP);
            if (c == 0) continue;
            const int yp = 14 * a + b;
            const bool kept = c == 1 || yp < YS::KSPLIT || yp >= YS::KSPLIT + YS::KSKIP;
            const int jr = (14 * a + 13 < YS::KSPLIT) ? yp : ((14 * a >= YS::KSPLIT + YS::KSKIP) ? yp - YS::KSKIP : (yp < YS::KSPLIT ? yp : yp - YS::KSKIP));
            cplx val = xo[a];
            bool store = act && kept;
            if (PAIR) {
              // even lane holds u'[y'], odd lane v'[y']: rows below the gap get u' + W^{-y'} v' (stored by the even lane),
              // rows above it u' - W^{-y'} v' (stored by the odd lane)
              const double2 tw = c_w252[yp];
              const cplx mine = role ? cmul_s<-1>(val, tw.x, tw.y) : val;
              cplx other;
              other.x = __shfl_xor_sync(0xffffffffu, mine.x, 16);
              other.y = __shfl_xor_sync(0xffffffffu, mine.y, 16);
              const bool low = yp < YS::KSPLIT;
              val = role ? make_double2(other.x - mine.x, other.y - mine.y) : make_double2(mine.x + other.x, mine.y + other.y);
              store = store && (low ? role == 0 : role == 1);
            }
            if (store) wz[(size_t)jr * np0] = val;
          }
        }
      }
    }
    if (OP == OP_DENSITY) {
      __syncthreads();
      double* rz = rho_part + (size_t)z * np01 + (size_t)xb * YS::COLS;
      const int ncol = min(YS::COLS, np0 - xb * YS::COLS);
      for (int i = tid; i < np1 * YS::COLS; i += 512) {
        const int y = i / YS::COLS, c = i - y * YS::COLS;
        if (c < ncol) rz[(size_t)y * np0 + c] += acc[i];
      }
    }
  }
  tmem_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc512(tbase);
}

// ------------------------------------------------------------------------------------------------ fused plane kernel, 126 x 126
// k_plane_f<OP, FS>: the WHOLE xy stage of a 126 x 126 plane (si54p) in one kernel, with no intermediate in HBM: the plane
// (254 KB) does not fit shared memory, its kept rows do (58 x 127 x 16 B = 118 KB), and the y direction needs no more than
// that because its columns live in tensor memory (the passes of k_ycols_t, reading the kept rows from shared memory instead
// of `w`).  Per unit: TMA bulk copy of the plane row (one unit ahead) -> scatter -> pruned x-DIT (compiled passes of
// plane_static.cuh, all 16 warps) -> y round trip in TMEM (all 16 warps, four per lane quarter) -> pruned x-DIF -> gather.
// Single-buffered: the two directions alternate (k_plane_t overlaps them on two buffers; 2 x 118 KB do not fit).
// grid (np2, G) persistent over the units gy, gy+G, ...; block 512.
template <int NP0_, int NP1_, int XSPLIT_, int YSPLIT_> struct FShape {
  static constexpr int NP0 = NP0_, NP1 = NP1_, XSPLIT = XSPLIT_, XSKIP = NP0_ - 2 * XSPLIT_, YSPLIT = YSPLIT_, YSKIP = NP1_ - 2 * YSPLIT_;
  static constexpr int NKEEP = 2 * YSPLIT_, PITCH = NP0_ | 1;
  static_assert(NP1_ == 126 && NP0_ <= 126, "thread-per-column y passes are written for 126 = 9 x 14, one plane per 128 TMEM lanes");
};
typedef FShape<126, 126, 29, 29> FsSi54p;

template <class FS> QB200_HD constexpr size_t plane_f_smem(int nvec, int nzero)
{
  constexpr FftDesc FX = make_fft_desc(FS::NP0);
  return (size_t)((FX.twsize + 7) & ~7) * 16 + (size_t)FS::NKEEP * FS::PITCH * 16 + (size_t)((nvec + 7) & ~7) * 16 +
         (size_t)((nvec + 7) & ~7) * 2 + (size_t)((nzero + 7) & ~7) * 2;
}

template <int OP, class FS>
__global__ void __launch_bounds__(512, 1) k_plane_f(const __grid_constant__ DevPlan P, cplx* __restrict__ zt, const double* __restrict__ v,
                                                    double* __restrict__ rho_part, const double* __restrict__ fac, int nunits)
{
  static_assert(OP == OP_HPSI || OP == OP_DENSITY, "k_plane_f: H psi and density only");
  constexpr FftDesc FX = make_fft_desc(FS::NP0);
  constexpr int np0 = FS::NP0, np1 = FS::NP1, np01 = np0 * np1, pitch = FS::PITCH, NK = FS::NKEEP, NT = 512;
  constexpr unsigned MASK = zmask(9, 14, FS::YSPLIT, FS::YSKIP);
  constexpr int MW = 4;
  extern __shared__ __align__(128) unsigned char smraw[];
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(8) uint64_t mbar;
  const int nvec = P.nvec, nvp = (nvec + 7) & ~7, nzero = P.ntzero;
  cplx* tw0 = reinterpret_cast<cplx*>(smraw);
  cplx* A = tw0 + ((FX.twsize + 7) & ~7);
  cplx* stg = A + NK * pitch;
  unsigned short* tpos = reinterpret_cast<unsigned short*>(stg + nvp);
  unsigned short* tzero = tpos + nvp;
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
  const int z = blockIdx.x, G = gridDim.y;
  const size_t N = (size_t)np01 * P.np2;
  if (warp == 0) tmem_alloc512(&tmem_slot);
  if (tid == 32) mbar_init(&mbar, 1);
  for (int i = tid; i < FX.twsize; i += NT) tw0[i] = P.tw0p[i];
  for (int i = tid; i < nvec; i += NT) tpos[i] = P.tpos[i];
  for (int i = tid; i < nzero; i += NT) tzero[i] = P.tzero[i];
  for (int i = tid; i < NK * pitch; i += NT) A[i] = make_double2(0.0, 0.0);
  tmem_fence_before();
  __syncthreads();
  tmem_fence_after();
  const uint32_t tbase = tmem_slot;
  const int q = warp & 3, m = warp >> 2;
  const uint32_t t0 = tbase + ((uint32_t)(q * 32) << 16);
  const int blo = (14 * m) / MW, bhi = (14 * (m + 1)) / MW, klo = (9 * m) / MW, khi = (9 * (m + 1)) / MW;
  const int cl = 32 * q + lane;
  const bool act = cl < np0;
  const int xc = min(cl, np0 - 1);
  const double* vz = v + (size_t)z * np01 + xc;
  double* rz = rho_part + (size_t)blockIdx.y * N + (size_t)z * np01 + xc;
  cplx* Ax = A + xc;
  auto csync = []() { __syncthreads(); };
  auto next_unit = [&](int u) {
    u += G;
    if (OP == OP_DENSITY) while (u < nunits && !fac_active(P, fac, u)) u += G;
    return u;
  };
  const uint32_t row_bytes = (uint32_t)nvec * 16u;
  uint32_t sphase = 0;
  int unit = next_unit((int)blockIdx.y - G);
  if (tid == 0 && unit < nunits) {
    mbar_expect_tx(&mbar, row_bytes);
    bulk_g2s(stg, zt + ((size_t)unit * P.np2 + z) * nvec, row_bytes, &mbar);
  }
  for (; unit < nunits;) {
    const int nxt = next_unit(unit);
    double facu = 0.0, facv = 0.0;
    if (OP == OP_DENSITY) { facu = fac_first(fac, unit); facv = fac_second(P, fac, unit); }
    cplx* ztrow = zt + ((size_t)unit * P.np2 + z) * nvec;
    mbar_wait(&mbar, sphase);
    sphase ^= 1u;
    __syncthreads();                               // everybody has left the previous unit (kept rows, TMEM slots)
    for (int j = tid; j < nzero; j += NT) A[tzero[j]] = make_double2(0.0, 0.0);
    for (int j = tid; j < nvec; j += NT) A[tpos[j]] = stg[j];
    __syncthreads();
    if (tid == 0 && nxt < nunits) {
      mbar_expect_tx(&mbar, row_bytes);
      bulk_g2s(stg, zt + ((size_t)nxt * P.np2 + z) * nvec, row_bytes, &mbar);
    }
    // x direction: kept rows, digit-reversed (zeros outside the sphere's h range) -> natural
    dit_s<+1, np0, 1, NK, DenseRowsW<pitch>, FS::XSPLIT, FS::XSKIP, true, false, FX.nf - 1>(tid, NT, A, tw0, csync);
    __syncthreads();
    // y pass 1: 9-point transforms over a of the kept rows y = 14a + b, twiddle -> slots (b, .)
#pragma unroll 1
    for (int b = blo; b < bhi; b++) {
      cplx xin[9];
#pragma unroll
      for (int a = 0; a < 9; a++) {
        const int c = zclass(a, 14, FS::YSPLIT, FS::YSKIP);
        if (c == 0) continue;
        const int yp = 14 * a + b;
        const bool kept = c == 1 || yp < FS::YSPLIT || yp >= FS::YSPLIT + FS::YSKIP;
        const int jr = (14 * a + 13 < FS::YSPLIT) ? yp : ((14 * a >= FS::YSPLIT + FS::YSKIP) ? yp - FS::YSKIP : (yp < FS::YSPLIT ? yp : yp - FS::YSKIP));
        xin[a] = kept ? Ax[jr * pitch] : make_double2(0.0, 0.0);
      }
      DftM<9, +1, MASK>::run(xin);
      if (b != 0) {
#pragma unroll
        for (int k1 = 1; k1 < 9; k1++) { const double2 tw = c_w126[9 * b + k1]; xin[k1] = cmul_s<+1>(xin[k1], tw.x, tw.y); }
      }
      Tmem<8>::st(t0 + 36 * b, xin);
      Tmem<1>::st(t0 + 36 * b + 32, xin + 8);
    }
    tmem_wait_st();
    tmem_fence_before();
    bar_sync_n(1 + q, 32 * MW);
    tmem_fence_after();
    // y pass 2: 14-point transforms over b -> psi(x, y = k1 + 9 k2, z); pointwise work; way back
#pragma unroll 1
    for (int k1 = klo; k1 < khi; k1++) {
      cplx t[14];
      Tmem<1, 14>::ld(t, t0 + 4 * k1, 36);
      Dft<14, +1>::run(t);
      if (OP == OP_HPSI) {
#pragma unroll
        for (int k2 = 0; k2 < 14; k2++) {
          const double vv = __ldg(vz + (size_t)(k1 + 9 * k2) * np0);
          t[k2].x *= vv;
          t[k2].y *= vv;
        }
        Dft<14, -1>::run(t);
#pragma unroll
        for (int b = 1; b < 14; b++) { const double2 tw = c_w126[9 * b + k1]; t[b] = cmul_s<-1>(t[b], tw.x, tw.y); }
        Tmem<1, 14>::st(t0 + 4 * k1, t, 36);
      } else if (act) {
#pragma unroll
        for (int k2 = 0; k2 < 14; k2++) {
          const double val = facu * t[k2].x * t[k2].x + facv * t[k2].y * t[k2].y;
          asm volatile("red.global.add.f64 [%0], %1;" ::"l"(rz + (size_t)(k1 + 9 * k2) * np0), "d"(val) : "memory");
        }
      }
    }
    if (OP == OP_HPSI) {
      tmem_wait_st();
      tmem_fence_before();
      bar_sync_n(1 + q, 32 * MW);
      tmem_fence_after();
      // y pass 3: 9-point transforms over k1 -> the kept rows y = 14a + b
#pragma unroll 1
      for (int b = blo; b < bhi; b++) {
        cplx xo[9];
        Tmem<8>::ld(xo, t0 + 36 * b);
        Tmem<1>::ld(xo + 8, t0 + 36 * b + 32);
        Dft<9, -1>::run(xo);
        if (act) {
#pragma unroll
          for (int a = 0; a < 9; a++) {
            const int c = zclass(a, 14, FS::YSPLIT, FS::YSKIP);
            if (c == 0) continue;
            const int yp = 14 * a + b;
            const bool kept = c == 1 || yp < FS::YSPLIT || yp >= FS::YSPLIT + FS::YSKIP;
            const int jr = (14 * a + 13 < FS::YSPLIT) ? yp : ((14 * a >= FS::YSPLIT + FS::YSKIP) ? yp - FS::YSKIP : (yp < FS::YSPLIT ? yp : yp - FS::YSKIP));
            if (kept) Ax[jr * pitch] = xo[a];
          }
        }
      }
      __syncthreads();
      dif_s<-1, np0, 1, NK, DenseRowsW<pitch>, FS::XSPLIT, FS::XSKIP, false, true, 0, FX.nf - 1>(tid, NT, A, tw0, csync);
      __syncthreads();
      for (int j = tid; j < nvec; j += NT) ztrow[j] = A[tpos[j]];
    }
    unit = nxt;
  }
  tmem_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc512(tbase);
}

// ------------------------------------------------------------------------------------------------ host side
static int ycols_t_shape(const qb200_plan* p)
{
  const DevPlan& d = p->d;
  if (d.np0 == YtSi54p::NP0 && d.np1 == YtSi54p::NP1 && d.ksplit == YtSi54p::YSPLIT && d.nkeep == YtSi54p::NKEEP) return 1;
  if (d.np0 == YtAu992::NP0 && d.np1 == YtAu992::NP1 && d.ksplit == YtAu992::YSPLIT && d.nkeep == YtAu992::NKEEP) return 2;
  return 0;
}

int ycols_t_setup(qb200_plan* p)
{
  p->ycols_t = 0;
  if (const char* e = getenv("QB200_YCOLS_T")) if (e[0] == '0') return QB200_OK;
  if (const char* e = getenv("QB200_NO_STATIC")) if (e[0] == '1') return QB200_OK;
  if (p->fused || !p->split2) return QB200_OK;
  const int shape = ycols_t_shape(p);
  if (!shape) return QB200_OK;
  const long double twopi = 6.283185307179586476925286766559005768L;
  double t126[2 * 14 * 9], t252[2 * 126];
  for (int b = 0; b < 14; b++)
    for (int k1 = 0; k1 < 9; k1++) {
      const int e = (b * k1) % 126;
      t126[2 * (9 * b + k1)] = (double)cosl(twopi * e / 126);
      t126[2 * (9 * b + k1) + 1] = (double)sinl(twopi * e / 126);
    }
  for (int y = 0; y < 126; y++) { t252[2 * y] = (double)cosl(twopi * y / 252); t252[2 * y + 1] = (double)sinl(twopi * y / 252); }
  QB_CUDA(cudaMemcpyToSymbol(c_w126, t126, sizeof(t126)));
  QB_CUDA(cudaMemcpyToSymbol(c_w252, t252, sizeof(t252)));
  QB_CUDA(cudaFuncSetAttribute(k_ycols_t<OP_DENSITY, YtSi54p>, cudaFuncAttributeMaxDynamicSharedMemorySize, YtSi54p::NP1 * YtSi54p::COLS * 8));
  QB_CUDA(cudaFuncSetAttribute(k_ycols_t<OP_DENSITY, YtAu992>, cudaFuncAttributeMaxDynamicSharedMemorySize, YtAu992::NP1 * YtAu992::COLS * 8));
  p->ycols_t = shape;
  return QB200_OK;
}

// fused 126 x 126 plane kernel (k_plane_f): QB200_PLANE_F=0 keeps the split path
bool plane_f_wanted(const qb200_plan* p, int hmax)
{
  if (const char* e = getenv("QB200_PLANE_F")) if (e[0] == '0') return false;
  if (const char* e = getenv("QB200_NO_STATIC")) if (e[0] == '1') return false;
  if (const char* e = getenv("QB200_FORCE_SPLIT")) if (e[0] == '1') return false;
  const DevPlan& d = p->d;
  typedef FsSi54p F;
  return !p->fused && d.np0 == F::NP0 && d.np1 == F::NP1 && d.ksplit == F::YSPLIT && d.nkeep == F::NKEEP && hmax < F::XSPLIT && d.nvec <= 65535;
}
int plane_f_pitch() { return FsSi54p::PITCH; }
void plane_f_xrange(int* xsplit, int* xskip) { *xsplit = FsSi54p::XSPLIT; *xskip = FsSi54p::XSKIP; }
int plane_f_setup(qb200_plan* p)
{
  p->plane_f = false;
  p->smem_plane_f = plane_f_smem<FsSi54p>(p->d.nvec, p->d.ntzero);
  if (p->smem_plane_f + 64 > (size_t)p->max_smem || !p->ycols_t) return QB200_OK;      // (the constants are set by ycols_t_setup)
  QB_CUDA((cudaFuncSetAttribute(k_plane_f<OP_HPSI, FsSi54p>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p->smem_plane_f)));
  QB_CUDA((cudaFuncSetAttribute(k_plane_f<OP_DENSITY, FsSi54p>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p->smem_plane_f)));
  p->plane_f = true;
  return QB200_OK;
}
int launch_plane_f(qb200_plan* p, int op, dim3 grid, const double* v, const double* fac, int nunits)
{
  if (op == OP_HPSI) k_plane_f<OP_HPSI, FsSi54p><<<grid, 512, p->smem_plane_f, p->stream>>>(p->d, (cplx*)p->zt, v, p->rho_part, fac, nunits);
  else k_plane_f<OP_DENSITY, FsSi54p><<<grid, 512, p->smem_plane_f, p->stream>>>(p->d, (cplx*)p->zt, v, p->rho_part, fac, nunits);
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "k_plane_f launch", __FILE__, __LINE__);
  return QB200_OK;
}

int launch_ycols_t(qb200_plan* p, int op, const double* v, const double* fac, int nunits, int zero_imag)
{
  const DevPlan& d = p->d;
  cplx* w = (cplx*)p->w;
  const int nslots = d.np2 * (p->ycols_t == 2 ? YtAu992::NXB : YtSi54p::NXB);
  const int grid = std::min(p->nsm, nslots);
  if (p->ycols_t == 1) {
    if (op == OP_HPSI) k_ycols_t<OP_HPSI, YtSi54p><<<grid, 512, 0, p->stream>>>(d, w, v, p->rho_part, fac, nunits, zero_imag);
    else k_ycols_t<OP_DENSITY, YtSi54p><<<grid, 512, (size_t)YtSi54p::NP1 * YtSi54p::COLS * 8, p->stream>>>(d, w, v, p->rho_part, fac, nunits, zero_imag);
  } else {
    if (op == OP_HPSI) k_ycols_t<OP_HPSI, YtAu992><<<grid, 512, 0, p->stream>>>(d, w, v, p->rho_part, fac, nunits, zero_imag);
    else k_ycols_t<OP_DENSITY, YtAu992><<<grid, 512, (size_t)YtAu992::NP1 * YtAu992::COLS * 8, p->stream>>>(d, w, v, p->rho_part, fac, nunits, zero_imag);
  }
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "k_ycols_t launch", __FILE__, __LINE__);
  return QB200_OK;
}

}  // namespace qb200
