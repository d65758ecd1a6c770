// qball_b200/csrc/split_kernels.cuh
// Second-generation split xy stage, for planes that do not fit one SM's shared memory (Au992: 252 x 252 x 16 B = 1 MB).
// Same three phases as the plane-fused kernel -- x transform of the kept rows (FourierTransform.cc:772-819), y
// transform of every column with the pointwise operation in the middle, x transform back -- as separate kernels with
// the compact kept-rows intermediate w[unit][z][jr][x] in HBM, rebuilt the way the z-column kernels were
// (zcol_kernels.cuh): persistent CTAs per (row block | column block, plane) looping over the units of a batch, every
// bulk global read an asynchronous copy issued one unit ahead, group-engine transforms without un-permute passes
// (scatter/gather/pointwise work absorb the digit reversal), pruned first/last y passes, and the density accumulated in
// shared memory across the units of a batch (one read-modify-write of rho per CTA and batch instead of one per state).
#pragma once
#include "qb200_internal.h"
#include "fft_group.cuh"
#include "zcol_kernels.cuh"
#include "plane_static.cuh"

namespace qb200 {

// SH selects the transform engine: DynSplit = run-time shapes (group engine, fft_group.cuh); SplitShape<...> = one grid
// shape known at compile time (pass_s of plane_static.cuh: immediate addressing, constant divisions, butterflies that
// know which inputs are zero / which outputs are never used).  Data flow and tables are identical.
struct DynSplit { static constexpr bool STATIC = false; };
template <int NP0_, int NP1_, int XSPLIT_, int XSKIP_, int YSPLIT_, int YSKIP_, int XB_, int ROWB_>
struct SplitShape {
  static constexpr bool STATIC = true;
  static constexpr int NP0 = NP0_, NP1 = NP1_, XSPLIT = XSPLIT_, XSKIP = XSKIP_, YSPLIT = YSPLIT_, YSKIP = YSKIP_;
  static constexpr int XB = XB_, ROWB = ROWB_, PITCH0 = NP0_ | 1, YPITCH = XB_ | 1, NKEEP = NP1_ - YSKIP_;
};
template <int PITCH> struct DenseRows {
  static __device__ __forceinline__ int off(int line) { return line * PITCH; }
};

// ------------------------------------------------------------------------------------------------ x rows
// grid (ceil(nkeep/rowb), np2, G), block 256.  DIR=+1: zt -> w (scatter to digit-reversed x, DIT, natural x out);
// DIR=-1: w -> zt (DIF, gather from digit-reversed x).
// smem: tw[f0.twsize] | rows[nbuf][rowb*pitch0] | stage[smax] (DIR=+1) | pos[smax] ints | src[smax] ints
template <int DIR, class SH>
__global__ void __launch_bounds__(256, 2) k_xrows2(const __grid_constant__ DevPlan P, cplx* __restrict__ zt, cplx* __restrict__ w,
                                                   int rowb, int smax, int nunits)
{
  extern __shared__ __align__(16) unsigned char smraw[];
  const int np0 = P.np0, pitch = P.pitch0;
  cplx* tw = reinterpret_cast<cplx*>(smraw);
  cplx* rows0 = tw + P.f0.twsize;
  const int nbuf = DIR > 0 ? 1 : 2;
  cplx* stage = rows0 + (size_t)nbuf * rowb * pitch;
  int* pos = reinterpret_cast<int*>(stage + (DIR > 0 ? smax : 0));
  int* src = pos + smax;
  const int tid = threadIdx.x, nthr = blockDim.x, G = gridDim.z;
  const int jr0 = blockIdx.x * rowb, jr1 = min(jr0 + rowb, P.nkeep), nr = jr1 - jr0;
  const int z = blockIdx.y;
  const int i0 = P.keeprowstart[jr0], cnt = P.keeprowstart[jr1] - i0;
  for (int i = tid; i < P.f0.twsize; i += nthr) tw[i] = P.tw0p[i];
  for (int i = tid; i < cnt; i += nthr) {
    const int iv = P.keepcols[i0 + i];
    src[i] = iv;
    pos[i] = (P.xs_jr[i0 + i] - jr0) * pitch + P.xs_x[i0 + i];
  }
  __syncthreads();
  const Grp g = { tid, nthr, 0 };
  const LineMap lm = { pitch, 1 << 30, 0 };
  const Keep nokeep = { 0, 0 };
  const FastDiv dx(np0);
  int unit = blockIdx.z;
  if (DIR > 0) {
    auto issue = [&](int u) {
      const cplx* ztrow = zt + ((size_t)u * P.np2 + z) * P.nvec;
      for (int i = tid; i < cnt; i += nthr) zc_cp16(stage + i, ztrow + src[i]);
      zc_commit();
    };
    if (unit < nunits) issue(unit);
    for (; unit < nunits; unit += G) {
      __syncthreads();                                   // previous write-out done with `rows`
      for (int i = tid; i < nr * pitch; i += nthr) rows0[i] = make_double2(0.0, 0.0);
      zc_wait_all();
      __syncthreads();
      for (int i = tid; i < cnt; i += nthr) rows0[pos[i]] = stage[i];
      __syncthreads();
      if (unit + G < nunits) issue(unit + G);
      if constexpr (SH::STATIC) {
        constexpr FftDesc FX = make_fft_desc(SH::NP0);
        dit_s<+1, SH::NP0, 1, SH::ROWB, DenseRows<SH::PITCH0>, SH::XSPLIT, SH::XSKIP, true, false, FX.nf - 1>(tid, nthr, rows0, tw, [] { __syncthreads(); }, nr);
      } else {
        fft_block_dit<+1>(g, rows0, nr, rowb, lm, 1, P.f0, tw, P.f0.nf - 1, false, nokeep);
      }
      __syncthreads();
      cplx* wz = w + (((size_t)unit * P.np2 + z) * P.nkeep + jr0) * np0;
      for (int e = tid; e < nr * np0; e += nthr) {
        int x;
        const int r = dx.div(e, x);
        wz[e] = rows0[r * pitch + x];
      }
    }
  } else {
    auto issue = [&](int u, int buf) {
      const cplx* wz = w + (((size_t)u * P.np2 + z) * P.nkeep + jr0) * np0;
      cplx* dst = rows0 + (size_t)buf * rowb * pitch;
      for (int e = tid; e < nr * np0; e += nthr) {
        int x;
        const int r = dx.div(e, x);
        zc_cp16(dst + r * pitch + x, wz + e);
      }
      zc_commit();
    };
    if (unit < nunits) issue(unit, 0);
    for (int buf = 0; unit < nunits; unit += G, buf ^= 1) {
      zc_wait_all();
      __syncthreads();
      if (unit + G < nunits) issue(unit + G, buf ^ 1);
      cplx* rows = rows0 + (size_t)buf * rowb * pitch;
      if constexpr (SH::STATIC) {
        constexpr FftDesc FX = make_fft_desc(SH::NP0);
        dif_s<-1, SH::NP0, 1, SH::ROWB, DenseRows<SH::PITCH0>, SH::XSPLIT, SH::XSKIP, false, true, 0, FX.nf - 1>(tid, nthr, rows, tw, [] { __syncthreads(); }, nr);
      } else {
        fft_block_dif<-1>(g, rows, nr, rowb, lm, 1, P.f0, tw, 0, P.f0.nf, false, nokeep);
      }
      __syncthreads();
      cplx* ztrow = zt + ((size_t)unit * P.np2 + z) * P.nvec;
      for (int i = tid; i < cnt; i += nthr) ztrow[src[i]] = rows[pos[i]];
    }
  }
  zc_wait_all();
}

// ------------------------------------------------------------------------------------------------ y columns
// grid (ceil(np0/xb), np2, G), block 256; CTA (bx, z, gy) is persistent over the units gy, gy+G, ... of the batch.
// smem: tw[f1.twsize] | yq[np1] ints (16-byte padded) | tile[np1*pitch] | stage[nkeep*xb] | acc[np1*xb] doubles (DENSITY)
// tile element (position q, column xl) at tile[q*pitch + xl], pitch = xb | 1; position q holds natural y = yq[q]
// between the DIF and the DIT transform.
template <int OP, class SH>
__global__ void __launch_bounds__(256, 2) k_ycols2(const __grid_constant__ DevPlan P, cplx* __restrict__ w, const double* __restrict__ v,
                                                   cplx* __restrict__ f, double* __restrict__ rho_part,
                                                   const double* __restrict__ fac, int nunits, int zero_imag)
{
  extern __shared__ __align__(16) unsigned char smraw[];
  const int np0 = P.np0, np1 = P.np1, np01 = np0 * np1, xb = P.xb, pitch = xb | 1, nkeep = P.nkeep;
  cplx* tw = reinterpret_cast<cplx*>(smraw);
  int* yq = reinterpret_cast<int*>(tw + P.f1.twsize);
  cplx* tile = reinterpret_cast<cplx*>(yq + ((np1 + 3) & ~3));
  cplx* stage = tile + (size_t)np1 * pitch;
  double* acc = reinterpret_cast<double*>(stage + (size_t)nkeep * xb);
  const int tid = threadIdx.x, nthr = blockDim.x, G = gridDim.z;
  const int x0 = blockIdx.x * xb, nx = min(xb, np0 - x0);
  const int z = blockIdx.y;
  const size_t N = (size_t)np01 * P.np2;
  for (int i = tid; i < P.f1.twsize; i += nthr) tw[i] = P.tw1p[i];
  for (int i = tid; i < np1; i += nthr) yq[i] = P.yq[i];
  if (OP == OP_DENSITY) for (int i = tid; i < np1 * xb; i += nthr) acc[i] = 0.0;
  const Grp g = { tid, nthr, 0 };
  const LineMap lm = { 1, 1 << 30, 0 };
  const Keep kp = { P.ksplit, P.kskip };
  const bool prune = nkeep < np1;
  const FastDiv dnx(nx);
  auto next_unit = [&](int u) {
    u += G;
    if (OP == OP_DENSITY) while (u < nunits && !fac_active(P, fac, u)) u += G;
    return u;
  };
  auto issue = [&](int u) {
    const cplx* wz = w + ((size_t)u * P.np2 + z) * nkeep * np0 + x0;
    for (int e = tid; e < nkeep * nx; e += nthr) {
      int xl;
      const int jr = dnx.div(e, xl);
      zc_cp16(stage + jr * xb + xl, wz + (size_t)jr * np0 + xl);
    }
    zc_commit();
  };
  int unit = next_unit((int)blockIdx.z - G);
  if (OP != OP_FWD && unit < nunits) issue(unit);
  __syncthreads();
  for (; unit < nunits;) {
    const int nxt = next_unit(unit);
    if (OP != OP_FWD) {
      zc_wait_all();
      __syncthreads();                                   // stage landed; previous unit's write-out done with the tile
      for (int e = tid; e < nkeep * nx; e += nthr) {
        int xl;
        const int jr = dnx.div(e, xl);
        tile[(jr < P.ksplit ? jr : jr + P.kskip) * pitch + xl] = stage[jr * xb + xl];
      }
      __syncthreads();
      if (nxt < nunits) issue(nxt);
      if constexpr (SH::STATIC) {
        constexpr FftDesc FY = make_fft_desc(SH::NP1);
        dif_s<+1, SH::NP1, SH::YPITCH, SH::XB, ColsOff, SH::YSPLIT, SH::YSKIP, true, false, 0, FY.nf - 1>(tid, nthr, tile, tw, [] { __syncthreads(); });
      } else {
        fft_block_dif<+1>(g, tile, nx, xb, lm, pitch, P.f1, tw, 0, P.f1.nf, prune, kp);
      }
      __syncthreads();
    } else {
      __syncthreads();
    }
    // pointwise work in digit-reversed y order
    if (OP == OP_HPSI) {
      const double* vz = v + (size_t)z * np01 + x0;
      for (int e = tid; e < np1 * nx; e += nthr) {
        int xl;
        const int q = dnx.div(e, xl);
        const double vv = __ldg(vz + (size_t)yq[q] * np0 + xl);
        cplx t = tile[q * pitch + xl];
        t.x *= vv;
        t.y = zero_imag ? 0.0 : t.y * vv;
        tile[q * pitch + xl] = t;
      }
    } else if (OP == OP_DENSITY) {
      const double facu = fac_first(fac, unit), facv = fac_second(P, fac, unit);
      for (int e = tid; e < np1 * nx; e += nthr) {
        int xl;
        const int q = dnx.div(e, xl);
        const cplx t = tile[q * pitch + xl];
        acc[q * xb + xl] += facu * t.x * t.x + facv * t.y * t.y;     // same thread owns the same (q, xl) for every unit
      }
    } else if (OP == OP_BWD) {
      cplx* fz = f + (size_t)unit * N + (size_t)z * np01 + x0;
      for (int e = tid; e < np1 * nx; e += nthr) {
        int xl;
        const int q = dnx.div(e, xl);
        fz[(size_t)yq[q] * np0 + xl] = tile[q * pitch + xl];
      }
    } else {
      const cplx* fz = f + (size_t)unit * N + (size_t)z * np01 + x0;
      for (int e = tid; e < np1 * nx; e += nthr) {
        int xl;
        const int q = dnx.div(e, xl);
        tile[q * pitch + xl] = fz[(size_t)yq[q] * np0 + xl];
      }
    }
    if (OP == OP_HPSI || OP == OP_FWD) {
      __syncthreads();
      if constexpr (SH::STATIC) {
        constexpr FftDesc FY = make_fft_desc(SH::NP1);
        dit_s<-1, SH::NP1, SH::YPITCH, SH::XB, ColsOff, SH::YSPLIT, SH::YSKIP, false, true, FY.nf - 1>(tid, nthr, tile, tw, [] { __syncthreads(); });
      } else {
        fft_block_dit<-1>(g, tile, nx, xb, lm, pitch, P.f1, tw, P.f1.nf - 1, prune, kp);
      }
      __syncthreads();
      cplx* wz = w + ((size_t)unit * P.np2 + z) * nkeep * np0 + x0;
      for (int e = tid; e < nkeep * nx; e += nthr) {
        int xl;
        const int jr = dnx.div(e, xl);
        wz[(size_t)jr * np0 + xl] = tile[(jr < P.ksplit ? jr : jr + P.kskip) * pitch + xl];
      }
    }
    unit = nxt;
  }
  zc_wait_all();
  if (OP == OP_DENSITY) {
    __syncthreads();
    double* rz = rho_part + (size_t)blockIdx.z * N + (size_t)z * np01 + x0;
    for (int e = tid; e < np1 * nx; e += nthr) {
      int xl;
      const int q = dnx.div(e, xl);
      rz[(size_t)yq[q] * np0 + xl] += acc[q * xb + xl];
    }
  }
}

}  // namespace qb200
