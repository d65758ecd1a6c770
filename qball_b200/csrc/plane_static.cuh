// qball_b200/csrc/plane_static.cuh
// k_plane_s<OP, SH>: the plane-fused xy stage of plane_kernels.cuh specialised at compile time for one grid shape SH
// (grid lengths, kept ranges, thread geometry).  Same data flow and the same tables as the generic k_plane2; what the
// specialisation buys:
//   * every shared-memory address is base + immediate, task decomposition divides by constants, no __noinline__ calls;
//   * the pruned passes know WHICH butterfly inputs are zero (fft_radix_masked.cuh) and which outputs are never used,
//     so that arithmetic disappears (FourierTransform.cc:202, 772-819 prunes whole transforms; here the pruning
//     reaches inside the first/last butterflies of both directions).
// Shapes compiled in: see plane.cu (the reference's benchmark grids); everything else runs the generic kernel.
#pragma once
#include "qb200_internal.h"
#include "fft_group.cuh"
#include "fft_radix_masked.cuh"
#include "plane_kernels.cuh"

namespace qb200 {

template <int NP0_, int NP1_, int XSPLIT_, int XSKIP_, int YSPLIT_, int YSKIP_, int NGRP_, int GT_, int PITCH_ = (NP0_ | 1)>
struct PlaneShape {
  static constexpr int NP0 = NP0_, NP1 = NP1_, PITCH = PITCH_;
  static constexpr int XSPLIT = XSPLIT_, XSKIP = XSKIP_;     // non-zero x: [0,XSPLIT) and [XSPLIT+XSKIP, NP0)
  static constexpr int YSPLIT = YSPLIT_, YSKIP = YSKIP_;     // kept rows:  [0,YSPLIT) and [YSPLIT+YSKIP, NP1)
  static constexpr int NKEEP = NP1_ - YSKIP_;
  static constexpr int NGRP = NGRP_, GT = GT_, NTHR = NGRP_ * GT_;
  static_assert(NP0_ % QB200_BLOCK_LINES == 0, "specialised shapes need np0 to be a multiple of the column block width");
  static_assert(NGRP_ * QB200_BLOCK_LINES <= NP0_ && NGRP_ <= 15 && GT_ % 32 == 0, "bad group geometry");
};

// class of element k of a task whose element 0 has natural index u in [0,step): 0 never kept, 1 always, 2 depends on u
QB200_HD constexpr int zclass(int k, int step, int split, int skip)
{
  bool any = false, all = true;
  for (int u = 0; u < step; u++) {
    const int j = u + k * step;
    const bool kp = j < split || j >= split + skip;
    any = any || kp;
    all = all && kp;
  }
  return all ? 1 : (any ? 2 : 0);
}
QB200_HD constexpr unsigned zmask(int R, int step, int split, int skip)
{
  unsigned m = 0;
  for (int k = 0; k < R; k++) if (zclass(k, step, split, skip) != 0) m |= 1u << k;
  return m;
}

template <int PITCH, int SPLIT, int SKIP> struct RowsOff {
  static __device__ __forceinline__ int off(int line) { return (line < SPLIT ? line : line + SKIP) * PITCH; }
};
struct ColsOff {
  static __device__ __forceinline__ int off(int line) { return line; }
};

// natural index of the first element of segment seg (R adjacent positions) of the last pass of a length-N transform
template <int N, int R> __device__ __forceinline__ int revseg(int seg)
{
  constexpr FftDesc F = make_fft_desc(N);
  int rem = seg, nat = 0, mul = 1, div = N / R;
#pragma unroll
  for (int i = 0; i < F.nf - 1; i++) {
    div /= F.r[i];
    const int dig = rem / div;
    rem -= dig * div;
    nat += dig * mul;
    mul *= F.r[i];
  }
  return nat;
}

enum { Z_NONE = 0, Z_IN = 1, Z_OUT = 2 };

// one radix-R pass, everything but the data known at compile time.  tw: packed twiddle table of the direction.
template <int R, int S, bool DIT, int N, int LEN, int TWOFF, int ESTRIDE, int NLPAD, int ZMODE, int SPLIT, int SKIP, class LINEOFF>
__device__ __forceinline__ void pass_s(int tid, int nthr, cplx* base, const cplx* tw, int nlv = NLPAD)
{
  constexpr int M = LEN / R, NTASK = (N / R) * NLPAD;
  constexpr int STEP = (M == 1) ? N / R : M;   // natural-index distance of a task's elements in a pruned pass
  constexpr unsigned MASK = ZMODE == Z_IN ? zmask(R, STEP, SPLIT, SKIP) : ((1u << R) - 1u);
  for (int task = tid; task < NTASK; task += nthr) {
    const int line = task % NLPAD, q = task / NLPAD;
    if (line >= nlv) continue;                  // (nlv == NLPAD unless the caller owns fewer lines: folds away)
    const int seg = q / M, t = q - seg * M;
    cplx* p = base + LINEOFF::off(line) + (seg * LEN + t) * ESTRIDE;
    const int u = (M == 1) ? revseg<N, R>(seg) : t;
    cplx x[R];
#pragma unroll
    for (int k = 0; k < R; k++) {
      if (ZMODE == Z_IN) {
        const int c = zclass(k, STEP, SPLIT, SKIP);
        if (c == 1) x[k] = p[k * M * ESTRIDE];
        else if (c == 2) {
          const int j = u + k * STEP;
          x[k] = (j < SPLIT || j >= SPLIT + SKIP) ? p[k * M * ESTRIDE] : make_double2(0.0, 0.0);
        }
      } else {
        x[k] = p[k * M * ESTRIDE];
      }
    }
    const cplx* twt = tw + TWOFF + t * (R - 1) - 1;
    if (DIT && M > 1) {
#pragma unroll
      for (int k = 1; k < R; k++) {
        if (!mask_bit(MASK, k)) continue;
        const cplx w = twt[k];
        x[k] = cmul_s<S>(x[k], w.x, w.y);
      }
    }
    DftM<R, S, MASK>::run(x);
    if (!DIT && M > 1) {
#pragma unroll
      for (int k = 1; k < R; k++) {
        const cplx w = twt[k];
        x[k] = cmul_s<S>(x[k], w.x, w.y);
      }
    }
#pragma unroll
    for (int k = 0; k < R; k++) {
      if (ZMODE == Z_OUT) {
        const int c = zclass(k, STEP, SPLIT, SKIP);
        if (c == 1) p[k * M * ESTRIDE] = x[k];
        else if (c == 2) {
          const int j = u + k * STEP;
          if (j < SPLIT || j >= SPLIT + SKIP) p[k * M * ESTRIDE] = x[k];
        }
      } else {
        p[k * M * ESTRIDE] = x[k];
      }
    }
  }
}

// passes first..last (inclusive, ascending) of a DIF transform of length N; `sync` between passes
template <int S, int N, int ESTRIDE, int NLPAD, class LINEOFF, int SPLIT, int SKIP, bool ZIN0, bool ZOUTLAST, int s, int last, class SYNC>
__device__ __forceinline__ void dif_s(int tid, int nthr, cplx* base, const cplx* tw, SYNC sync, int nlv = NLPAD)
{
  constexpr FftDesc F = make_fft_desc(N);
  constexpr int Z = (s == 0 && ZIN0) ? Z_IN : ((s == F.nf - 1 && ZOUTLAST) ? Z_OUT : Z_NONE);
  pass_s<F.r[s], S, false, N, F.len[s], F.twoff[s], ESTRIDE, NLPAD, Z, SPLIT, SKIP, LINEOFF>(tid, nthr, base, tw, nlv);
  if constexpr (s < last) {
    sync();
    dif_s<S, N, ESTRIDE, NLPAD, LINEOFF, SPLIT, SKIP, ZIN0, ZOUTLAST, s + 1, last, SYNC>(tid, nthr, base, tw, sync, nlv);
  }
}

// passes first..0 (descending) of a DIT transform of length N
template <int S, int N, int ESTRIDE, int NLPAD, class LINEOFF, int SPLIT, int SKIP, bool ZINFIRST, bool ZOUT0, int s, class SYNC>
__device__ __forceinline__ void dit_s(int tid, int nthr, cplx* base, const cplx* tw, SYNC sync, int nlv = NLPAD)
{
  constexpr FftDesc F = make_fft_desc(N);
  constexpr int Z = (s == F.nf - 1 && ZINFIRST) ? Z_IN : ((s == 0 && ZOUT0) ? Z_OUT : Z_NONE);
  pass_s<F.r[s], S, true, N, F.len[s], F.twoff[s], ESTRIDE, NLPAD, Z, SPLIT, SKIP, LINEOFF>(tid, nthr, base, tw, nlv);
  if constexpr (s > 0) {
    sync();
    dit_s<S, N, ESTRIDE, NLPAD, LINEOFF, SPLIT, SKIP, ZINFIRST, ZOUT0, s - 1, SYNC>(tid, nthr, base, tw, sync, nlv);
  }
}

// middle pass of the y direction on a block of 8 columns (cf. mid_pass in plane_kernels.cuh)
template <int OP, class SH>
__device__ __forceinline__ void mid_s(int tid, int nthr, cplx* blk, int nlines, const MidArgs& a)
{
  constexpr FftDesc F = make_fft_desc(SH::NP1);
  constexpr int R = F.r[F.nf - 1], NSEG = SH::NP1 / R, NTASK = NSEG * QB200_BLOCK_LINES;
  constexpr bool ONLY1 = F.nf == 1;
  constexpr unsigned MASK = ONLY1 ? zmask(R, 1, SH::YSPLIT, SH::YSKIP) : ((1u << R) - 1u);
  constexpr size_t YSTEP = (size_t)NSEG * SH::NP0;
  for (int task = tid; task < NTASK; task += nthr) {
    const int line = task & (QB200_BLOCK_LINES - 1);
    if ((SH::NP0 % QB200_BLOCK_LINES) != 0 && line >= nlines) continue;
    const int seg = task >> 3;
    cplx* p = blk + line + seg * R * SH::PITCH;
    const size_t g0 = (size_t)revseg<SH::NP1, R>(seg) * SH::NP0 + line;
    cplx x[R];
    if (OP == OP_FWD) {
      const cplx* fp = a.f + g0;
#pragma unroll
      for (int j = 0; j < R; j++) x[j] = fp[j * YSTEP];
    } else {
      double vv[R];
      if (OP == OP_HPSI) {
        const double* vp = a.v + g0;
#pragma unroll
        for (int j = 0; j < R; j++) vv[j] = (a.exp & 2) ? 1.0 : __ldg(vp + j * YSTEP);
      }
#pragma unroll
      for (int k = 0; k < R; k++) if (mask_bit(MASK, k)) x[k] = p[k * SH::PITCH];
      DftM<R, +1, MASK>::run(x);
      if (OP == OP_HPSI) {
#pragma unroll
        for (int j = 0; j < R; j++) { x[j].x *= vv[j]; x[j].y = a.zero_imag ? 0.0 : x[j].y * vv[j]; }
      } else if (OP == OP_DENSITY) {
        double* rp = a.rho + g0;
        // fire-and-forget reduction at the L2 (red.global.add.f64) instead of load / add / store: no load latency in the
        // pass and half the LSU instructions (xy stage 14.6 -> 14.3 ms).  Every address is owned by one thread of one CTA,
        // whose reductions to it are applied in program (= unit) order: deterministic.
#pragma unroll
        if (a.exp & 1) {          // timing experiment: no reductions (one dependent store keeps the arithmetic alive)
          double acc = 0.0;
#pragma unroll
          for (int j = 0; j < R; j++) acc += a.facu * x[j].x * x[j].x + a.facv * x[j].y * x[j].y;
          if (acc == 1.2345e300) rp[0] = acc;
        } else
#pragma unroll
        for (int j = 0; j < R; j++) {
          const double val = a.facu * x[j].x * x[j].x + a.facv * x[j].y * x[j].y;
          asm volatile("red.global.add.f64 [%0], %1;" ::"l"(rp + j * YSTEP), "d"(val) : "memory");
        }
      } else if (OP == OP_BWD) {
        cplx* fp = a.f + g0;
#pragma unroll
        for (int j = 0; j < R; j++) fp[j * YSTEP] = x[j];
      }
    }
    if (OP == OP_HPSI || OP == OP_FWD) {
      Dft<R, -1>::run(x);
#pragma unroll
      for (int k = 0; k < R; k++) if (mask_bit(MASK, k)) p[k * SH::PITCH] = x[k];
    }
  }
}

// grid (np2, G), block SH::NTHR; shared memory and tables exactly as k_plane2 (plane_kernels.cuh)
template <int OP, class SH>
__global__ void __launch_bounds__(SH::NTHR, 1) k_plane_s(const __grid_constant__ DevPlan P, cplx* __restrict__ zt, const double* __restrict__ v,
                                                         cplx* __restrict__ f, double* __restrict__ rho_part,
                                                         const double* __restrict__ fac, int nunits, int zero_imag)
{
  constexpr FftDesc FX = make_fft_desc(SH::NP0), FY = make_fft_desc(SH::NP1);
  constexpr int np0 = SH::NP0, np1 = SH::NP1, pitch = SH::PITCH, np01 = np0 * np1;
  extern __shared__ __align__(16) unsigned char smraw[];
  cplx* tw0 = reinterpret_cast<cplx*>(smraw);
  cplx* tw1 = tw0 + FX.twsize;
  int* colpos_s = reinterpret_cast<int*>(tw1 + FY.twsize + P.nyrev_c);
  cplx* pl = tw1 + FY.twsize + P.nyrev_c + P.ncolpos_c;
  const int nvec = P.nvec;
  const int z = blockIdx.x;
  const size_t N = (size_t)np01 * P.np2;
  const int tid = threadIdx.x;
  for (int i = tid; i < FX.twsize; i += SH::NTHR) tw0[i] = P.tw0p[i];
  for (int i = tid; i < FY.twsize; i += SH::NTHR) tw1[i] = P.tw1p[i];
  if (P.ncolpos_c) for (int i = tid; i < nvec; i += SH::NTHR) colpos_s[i] = P.colpos[i];
  const int* colpos = P.ncolpos_c ? colpos_s : P.colpos;
  const int gid = tid / SH::GT, gtid = tid - gid * SH::GT;
  const int gbar = 1 + gid;
  // one warp per group: a warp-level sync is enough (a named barrier also drains the warp's pending shared-memory stores)
  auto gsync = [gbar]() {
    if constexpr (SH::GT == 32) __syncwarp();
    else asm volatile("bar.sync %0, %1;" ::"r"(gbar), "n"(SH::GT) : "memory");
  };
  auto csync = []() { __syncthreads(); };
  constexpr int hi0 = (SH::YSPLIT + SH::YSKIP) * pitch;
  constexpr int nlo = SH::YSPLIT * pitch, nhi = (np1 - SH::YSPLIT - SH::YSKIP) * pitch;
  const int per = (OP == OP_FWD) ? 0 : P.stage_per;
  const int G = gridDim.y;
  typedef RowsOff<pitch, SH::YSPLIT, SH::YSKIP> ROWS;
  auto stage = [&](int unit) {
    const cplx* src = zt + ((size_t)unit * P.np2 + z) * nvec + gid * per;
    const int cnt = min(per, nvec - gid * per);
    cplx* dst = pl + SH::YSPLIT * pitch + gid * QB200_BLOCK_LINES;
    for (int j = gtid; j < cnt; j += SH::GT) cp_async16(dst + (j >> 3) * pitch + (j & 7), src + j);
  };
  auto next_unit = [&](int u) {
    u += G;
    if (OP == OP_DENSITY) while (u < nunits && !fac_active(P, fac, u)) u += G;
    return u;
  };
  int unit = next_unit((int)blockIdx.y - G);
  if (per && unit < nunits) stage(unit);
  for (; unit < nunits;) {
    const int nxt = next_unit(unit);
    double facu = 0.0, facv = 0.0;
    if (OP == OP_DENSITY) { facu = fac_first(fac, unit); facv = fac_second(P, fac, unit); }
    cplx* ztrow = zt + ((size_t)unit * P.np2 + z) * nvec;
    if (per) cp_async_wait_all();
    __syncthreads();
    if (OP != OP_FWD) {
      if (!(P.exp & 4)) {
      for (int i = tid; i < nlo; i += SH::NTHR) pl[i] = make_double2(0.0, 0.0);
      for (int i = tid; i < nhi; i += SH::NTHR) pl[hi0 + i] = make_double2(0.0, 0.0);
      }
      __syncthreads();
      if (P.exp & 4) {
      } else if (per) {
        const FastDiv dp(per);
        for (int i = tid; i < nvec; i += SH::NTHR) {
          int j;
          const int sh = dp.div(i, j);
          pl[colpos[i]] = pl[(SH::YSPLIT + (j >> 3)) * pitch + sh * QB200_BLOCK_LINES + (j & 7)];
        }
      } else {
        constexpr int U = 8;
        for (int i0 = tid; i0 < nvec; i0 += U * SH::NTHR) {
          cplx val[U];
#pragma unroll
          for (int u = 0; u < U; u++) val[u] = ztrow[min(i0 + u * SH::NTHR, nvec - 1)];
#pragma unroll
          for (int u = 0; u < U; u++) { const int i = i0 + u * SH::NTHR; if (i < nvec) pl[colpos[i]] = val[u]; }
        }
      }
      __syncthreads();
      // x direction, whole CTA: kept rows, digit-reversed (zeros outside the sphere's h range) -> natural
      dit_s<+1, np0, 1, SH::NKEEP, ROWS, SH::XSPLIT, SH::XSKIP, true, false, FX.nf - 1>(tid, SH::NTHR, pl, tw0, csync);
      __syncthreads();
    }
    // y direction: blocks of 8 columns, one group each
    for (int b = gid; b * QB200_BLOCK_LINES < np0; b += SH::NGRP) {
      const int c0 = b * QB200_BLOCK_LINES;
      const int nc = min(QB200_BLOCK_LINES, np0 - c0);
      cplx* blk = pl + c0;
      if constexpr (OP != OP_FWD && FY.nf > 1) {
        dif_s<+1, np1, pitch, QB200_BLOCK_LINES, ColsOff, SH::YSPLIT, SH::YSKIP, true, false, 0, FY.nf - 2>(gtid, SH::GT, blk, tw1, gsync);
        gsync();
      }
      MidArgs a;
      a.v = v + (size_t)z * np01 + c0;
      a.rho = rho_part + (size_t)blockIdx.y * N + (size_t)z * np01 + c0;
      a.f = f + (size_t)unit * N + (size_t)z * np01 + c0;
      a.facu = facu; a.facv = facv; a.np0 = np0; a.zero_imag = zero_imag; a.exp = P.exp;
      mid_s<OP, SH>(gtid, SH::GT, blk, nc, a);
      if constexpr ((OP == OP_HPSI || OP == OP_FWD) && FY.nf > 1) {
        gsync();
        dit_s<-1, np1, pitch, QB200_BLOCK_LINES, ColsOff, SH::YSPLIT, SH::YSKIP, false, true, FY.nf - 2>(gtid, SH::GT, blk, tw1, gsync);
      }
      if (per && b == gid && nxt < nunits) {
        gsync();
        stage(nxt);
      }
    }
    if (OP == OP_HPSI || OP == OP_FWD) {
      __syncthreads();
      dif_s<-1, np0, 1, SH::NKEEP, ROWS, SH::XSPLIT, SH::XSKIP, false, true, 0, FX.nf - 1>(tid, SH::NTHR, pl, tw0, csync);
      __syncthreads();
      for (int i = tid; i < nvec; i += SH::NTHR) ztrow[i] = pl[colpos[i]];
    }
    unit = nxt;
  }
}


// ------------------------------------------------------------------------------------------------ warp-owned variant
// k_plane_w<OP, SH>: k_plane_s with one WARP per group (SH::GT == 32) and the x phase made warp-local as well: warp g owns
// RBX kept rows (low half for g < NGRP/2, high half above) and, through a table sorted by owner (P.wown: staged address |
// plane position << 16, P.wown_iv: column index), zero-fills, scatters, x-transforms and finally gathers ITS rows without
// any CTA barrier -- three CTA barriers per unit remain (staged data landed, x -> y, y -> x) instead of eight, and the
// warps drift out of phase so that one warp's shared-memory traffic overlaps another's FP64 butterflies in the x phase
// too.  The row pitch is 2 (mod 8) 16-byte slots, which keeps the 4-row x passes bank-conflict free.
template <int PITCH> struct DenseRowsW {
  static __device__ __forceinline__ int off(int line) { return line * PITCH; }
};

template <int OP, class SH>
__global__ void __launch_bounds__(SH::NTHR, 1) k_plane_w(const __grid_constant__ DevPlan P, cplx* __restrict__ zt, const double* __restrict__ v,
                                                         cplx* __restrict__ f, double* __restrict__ rho_part,
                                                         const double* __restrict__ fac, int nunits, int zero_imag)
{
  static_assert(SH::GT == 32 && SH::NGRP % 2 == 0 && SH::NP1 - SH::YSPLIT - SH::YSKIP == SH::YSPLIT, "warp-owned geometry");
  static_assert(SH::NGRP * QB200_BLOCK_LINES == SH::NP0, "one column block per warp");
  constexpr FftDesc FX = make_fft_desc(SH::NP0), FY = make_fft_desc(SH::NP1);
  constexpr int np0 = SH::NP0, np1 = SH::NP1, pitch = SH::PITCH, np01 = np0 * np1;
  constexpr int HALF = SH::NGRP / 2, RBX = (SH::YSPLIT + HALF - 1) / HALF;       // kept rows per warp
  extern __shared__ __align__(16) unsigned char smraw[];
  cplx* tw0 = reinterpret_cast<cplx*>(smraw);
  cplx* tw1 = tw0 + FX.twsize;
  int* own_s = reinterpret_cast<int*>(tw1 + FY.twsize + P.nyrev_c);
  int* owniv_s = own_s + 4 * P.ncolpos_c;
  cplx* pl = tw1 + FY.twsize + P.nyrev_c + 2 * P.ncolpos_c;
  const int nvec = P.nvec;
  const int z = blockIdx.x;
  const size_t N = (size_t)np01 * P.np2;
  const int tid = threadIdx.x;
  for (int i = tid; i < FX.twsize; i += SH::NTHR) tw0[i] = P.tw0p[i];
  for (int i = tid; i < FY.twsize; i += SH::NTHR) tw1[i] = P.tw1p[i];
  for (int i = tid; i < nvec; i += SH::NTHR) { own_s[i] = P.wown[i]; owniv_s[i] = P.wown_iv[i]; }
  const int gid = tid >> 5, lane = tid & 31;
  auto wsync = []() { __syncwarp(); };
  // this warp's kept rows and its entries of the owner-sorted table
  const int rlo = (gid < HALF) ? gid * RBX : (gid - HALF) * RBX;
  const int nrow = max(0, min(RBX, SH::YSPLIT - rlo));
  const int row0 = (gid < HALF) ? rlo : SH::YSPLIT + SH::YSKIP + rlo;
  const int e0 = P.wown_start[gid], e1 = P.wown_start[gid + 1];
  cplx* myrows = pl + row0 * pitch;
  const int per = (OP == OP_FWD) ? 0 : P.stage_per;
  const int G = gridDim.y;
  auto stage = [&](int unit) {
    const cplx* src = zt + ((size_t)unit * P.np2 + z) * nvec + gid * per;
    const int cnt = min(per, nvec - gid * per);
    cplx* dst = pl + SH::YSPLIT * pitch + gid * QB200_BLOCK_LINES;
    for (int j = lane; j < cnt; j += 32) cp_async16(dst + (j >> 3) * pitch + (j & 7), src + j);
  };
  auto next_unit = [&](int u) {
    u += G;
    if (OP == OP_DENSITY) while (u < nunits && !fac_active(P, fac, u)) u += G;
    return u;
  };
  int unit = next_unit((int)blockIdx.y - G);
  if (per && unit < nunits) stage(unit);
  for (; unit < nunits;) {
    const int nxt = next_unit(unit);
    double facu = 0.0, facv = 0.0;
    if (OP == OP_DENSITY) { facu = fac_first(fac, unit); facv = fac_second(P, fac, unit); }
    cplx* ztrow = zt + ((size_t)unit * P.np2 + z) * nvec;
    if (per) cp_async_wait_all();
    __syncthreads();                 // tables / every warp's staged values visible; the plane is free
    if (OP != OP_FWD) {
      for (int i = lane; i < nrow * pitch; i += 32) myrows[i] = make_double2(0.0, 0.0);
      __syncwarp();
      if (per) {
        for (int e = e0 + lane; e < e1; e += 32) { const int w = own_s[e]; pl[w >> 16] = pl[w & 0xffff]; }
      } else {
        for (int e = e0 + lane; e < e1; e += 32) pl[own_s[e] >> 16] = ztrow[owniv_s[e]];
      }
      __syncwarp();
      // x direction: this warp's rows, digit-reversed (zeros outside the sphere's h range) -> natural
      dit_s<+1, np0, 1, RBX, DenseRowsW<pitch>, SH::XSPLIT, SH::XSKIP, true, false, FX.nf - 1>(lane, 32, myrows, tw0, wsync, nrow);
      __syncthreads();
    }
    // y direction: the warp's block of 8 columns
    {
      const int c0 = gid * QB200_BLOCK_LINES;
      cplx* blk = pl + c0;
      if constexpr (OP != OP_FWD && FY.nf > 1) {
        dif_s<+1, np1, pitch, QB200_BLOCK_LINES, ColsOff, SH::YSPLIT, SH::YSKIP, true, false, 0, FY.nf - 2>(lane, 32, blk, tw1, wsync);
        __syncwarp();
      }
      MidArgs a;
      a.v = v + (size_t)z * np01 + c0;
      a.rho = rho_part + (size_t)blockIdx.y * N + (size_t)z * np01 + c0;
      a.f = f + (size_t)unit * N + (size_t)z * np01 + c0;
      a.facu = facu; a.facv = facv; a.np0 = np0; a.zero_imag = zero_imag; a.exp = 0;
      mid_s<OP, SH>(lane, 32, blk, QB200_BLOCK_LINES, a);
      if constexpr ((OP == OP_HPSI || OP == OP_FWD) && FY.nf > 1) {
        __syncwarp();
        dit_s<-1, np1, pitch, QB200_BLOCK_LINES, ColsOff, SH::YSPLIT, SH::YSKIP, false, true, FY.nf - 2>(lane, 32, blk, tw1, wsync);
      }
      if (per && nxt < nunits) {
        __syncwarp();                // the warp is done with its block before the block's dead rows are overwritten
        stage(nxt);
      }
    }
    if (OP == OP_HPSI || OP == OP_FWD) {
      __syncthreads();
      dif_s<-1, np0, 1, RBX, DenseRowsW<pitch>, SH::XSPLIT, SH::XSKIP, false, true, 0, FX.nf - 1>(lane, 32, myrows, tw0, wsync, nrow);
      __syncwarp();
      for (int e = e0 + lane; e < e1; e += 32) ztrow[owniv_s[e]] = pl[own_s[e] >> 16];
    }
    unit = nxt;
  }
}

}  // namespace qb200
