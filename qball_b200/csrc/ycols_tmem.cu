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
      double facu = 0.0;
      if (OP == OP_DENSITY) { facu = fac[unit]; if (!(facu > 0.0)) continue; }
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
            acc[y * YS::COLS + cl] += facu * (t[k2].x * t[k2].x + t[k2].y * t[k2].y);    // this thread owns (y, cl) in every unit
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
