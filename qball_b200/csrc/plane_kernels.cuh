// qball_b200/csrc/plane_kernels.cuh
// k_plane2<OP>: the plane-fused xy stage (second generation).  One CTA holds one xy-plane of one unit in shared
// memory; the CTA is divided into groups of P.gthreads threads that own blocks of 8 rows (x direction) or 8 columns
// (y direction) and synchronise among themselves only (fft_group.cuh).  Per plane:
//
//   zero kept rows -> scatter the plane's nvec column values (digit-reversed x positions)            [CTA barrier]
//   x: DIT transform of the 2*ntrans0 kept rows (FourierTransform.cc:772-819), natural x out          [CTA barrier]
//   y, per block of 8 columns: first DIF pass loads only the kept rows (the others are zero) ... last DIF butterfly
//      + OP + first DIT butterfly in registers ("mid" pass) ... last DIT pass stores only the kept rows
//                                                                                                    [CTA barrier]
//   x: DIF transform of the kept rows, gather the nvec values from digit-reversed x positions.
//
//   OP_HPSI    psi(r) *= v(r)                  (SlaterDet::rs_mul_add, SlaterDet.cc:993-1031)
//   OP_DENSITY rho_part += fac*|psi(r)|^2       (SlaterDet::compute_density, SlaterDet.cc:919-921); no way back
//   OP_BWD     f = psi(r)                       (FourierTransform::backward)
//   OP_FWD     psi(r) = f, way back only        (FourierTransform::forward)
#pragma once
#include "qb200_internal.h"
#include "fft_group.cuh"

namespace qb200 {

struct MidArgs {
  const double* v;      // v + z*np01 + c0          (OP_HPSI)
  double* rho;          // rho_part plane + c0      (OP_DENSITY)
  cplx* f;              // f plane + c0             (OP_BWD / OP_FWD)
  double facu, facv;    // weights of |Re psi|^2 and |Im psi|^2 (equal unless the unit is a pair of real states)
  int np0;
  int zero_imag;
  int exp;
};

// The middle pass of the y direction on a block of columns: element (line, j) at base[line + j*estride].
// Task = (column, segment of R adjacent positions).  Position seg*R + j holds natural y = yrev[seg] + j*(n/R).
// only1 (n == R): the pass is also the first and the last one -> pruned loads/stores.
template <int R, int OP>
__device__ __noinline__ void mid_pass(Grp g, cplx* base, int nlines, int estride, int n, const int* __restrict__ yrev,
                                      MidArgs a, bool only1, Keep kp)
{
  const int nseg = n / R;
  const int ntask = nseg * QB200_BLOCK_LINES;
  const size_t ystep = (size_t)nseg * a.np0;
  for (int task = g.tid; task < ntask; task += g.nthr) {
    const int line = task & (QB200_BLOCK_LINES - 1);
    if (line >= nlines) continue;
    const int seg = task >> 3;
    cplx* p = base + line + seg * R * estride;
    const size_t g0 = (size_t)yrev[seg] * a.np0 + line;
    cplx x[R];
    if (OP == OP_FWD) {
      const cplx* fp = a.f + g0;
#pragma unroll
      for (int j = 0; j < R; j++) x[j] = fp[j * ystep];
    } else {
      double vv[R];
      if (OP == OP_HPSI) {
        const double* vp = a.v + g0;
#pragma unroll
        for (int j = 0; j < R; j++) vv[j] = __ldg(vp + j * ystep);
      } else if (OP == OP_DENSITY) {
        const double* rp = a.rho + g0;
#pragma unroll
        for (int j = 0; j < R; j++) vv[j] = rp[j * ystep];
      }
      if (only1) {
#pragma unroll
        for (int k = 0; k < R; k++) x[k] = kp.kept(k) ? p[k * estride] : make_double2(0.0, 0.0);
      } else {
#pragma unroll
        for (int k = 0; k < R; k++) x[k] = p[k * estride];
      }
      Dft<R, +1>::run(x);
      if (OP == OP_HPSI) {
#pragma unroll
        for (int j = 0; j < R; j++) { x[j].x *= vv[j]; x[j].y = a.zero_imag ? 0.0 : x[j].y * vv[j]; }
      } else if (OP == OP_DENSITY) {
        double* rp = a.rho + g0;
#pragma unroll
        for (int j = 0; j < R; j++) rp[j * ystep] = vv[j] + (a.facu * x[j].x * x[j].x + a.facv * x[j].y * x[j].y);
      } else if (OP == OP_BWD) {
        cplx* fp = a.f + g0;
#pragma unroll
        for (int j = 0; j < R; j++) fp[j * ystep] = x[j];
      }
    }
    if (OP == OP_HPSI || OP == OP_FWD) {
      Dft<R, -1>::run(x);
      if (only1) {
#pragma unroll
        for (int k = 0; k < R; k++) if (kp.kept(k)) p[k * estride] = x[k];
      } else {
#pragma unroll
        for (int k = 0; k < R; k++) p[k * estride] = x[k];
      }
    }
  }
}

template <int OP>
__device__ __forceinline__ void mid_pass_any(int r, Grp g, cplx* base, int nlines, int estride, int n,
                                             const int* __restrict__ yrev, MidArgs a, bool only1, Keep kp)
{
  switch (r) {
    case 16: mid_pass<16, OP>(g, base, nlines, estride, n, yrev, a, only1, kp); break;
    case 8: mid_pass<8, OP>(g, base, nlines, estride, n, yrev, a, only1, kp); break;
    case 4: mid_pass<4, OP>(g, base, nlines, estride, n, yrev, a, only1, kp); break;
    case 2: mid_pass<2, OP>(g, base, nlines, estride, n, yrev, a, only1, kp); break;
    case 9: mid_pass<9, OP>(g, base, nlines, estride, n, yrev, a, only1, kp); break;
    case 3: mid_pass<3, OP>(g, base, nlines, estride, n, yrev, a, only1, kp); break;
    case 5: mid_pass<5, OP>(g, base, nlines, estride, n, yrev, a, only1, kp); break;
    case 7: mid_pass<7, OP>(g, base, nlines, estride, n, yrev, a, only1, kp); break;
    case 11: mid_pass<11, OP>(g, base, nlines, estride, n, yrev, a, only1, kp); break;
    case 1: mid_pass<1, OP>(g, base, nlines, estride, n, yrev, a, only1, kp); break;
    default: break;
  }
}

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem)
{
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// grid (np2, G): CTA (z, gy) is persistent over the units gy, gy+G, ... of plane z; block P.plane_threads (a multiple
// of P.gthreads).  dynamic smem: twiddles (np0 + np1), yrev and colpos tables, the plane np1 x pitch0.
// While a unit is in its y phase, each group copies its share of the NEXT unit's column values (cp.async) into the
// rows of its first column block that are dead by then (not kept: zero before the y phase, not needed after it), so
// the next scatter runs out of shared memory and no global-memory latency is exposed between units.
template <int OP>
__global__ void __launch_bounds__(448, 1) k_plane2(const __grid_constant__ DevPlan P, cplx* __restrict__ zt, const double* __restrict__ v,
                                                    cplx* __restrict__ f, double* __restrict__ rho_part,
                                                    const double* __restrict__ fac, int nunits, int zero_imag)
{
  extern __shared__ __align__(16) unsigned char smraw[];
  cplx* tw0 = reinterpret_cast<cplx*>(smraw);
  cplx* tw1 = tw0 + P.f0.twsize;
  int* yrev = reinterpret_cast<int*>(tw1 + P.f1.twsize);
  int* colpos_s = reinterpret_cast<int*>(tw1 + P.f1.twsize + P.nyrev_c);
  cplx* pl = tw1 + P.f1.twsize + P.nyrev_c + P.ncolpos_c;
  const int np0 = P.np0, np1 = P.np1, pitch = P.pitch0, np01 = np0 * np1, nvec = P.nvec;
  const int z = blockIdx.x;
  const size_t N = (size_t)np01 * P.np2;
  for (int i = threadIdx.x; i < P.f0.twsize; i += blockDim.x) tw0[i] = P.tw0p[i];
  for (int i = threadIdx.x; i < P.f1.twsize; i += blockDim.x) tw1[i] = P.tw1p[i];
  const int rl = P.f1.r[P.f1.nf - 1];
  for (int i = threadIdx.x; i < np1 / rl; i += blockDim.x) yrev[i] = P.yrev[i];
  if (P.ncolpos_c) for (int i = threadIdx.x; i < nvec; i += blockDim.x) colpos_s[i] = P.colpos[i];
  const int* colpos = P.ncolpos_c ? colpos_s : P.colpos;
  const int gsz = P.gthreads, ngrp = blockDim.x / gsz, gid = threadIdx.x / gsz;
  const Grp g = { (int)threadIdx.x - gid * gsz, gsz, 1 + gid };
  const Grp cta = { (int)threadIdx.x, (int)blockDim.x, 0 };
  const Keep keepy = { P.ksplit, P.kskip };
  const Keep all = { 1 << 30, 0 };
  const LineMap rows = { pitch, P.ksplit, P.kskip };
  const LineMap colsmap = { 1, 1 << 30, 0 };
  const int hi0 = (P.ksplit + P.kskip) * pitch;   // first element of the upper kept rows
  const int nlo = P.ksplit * pitch, nhi = (np1 - P.ksplit - P.kskip) * pitch;
  const int per = (OP == OP_FWD) ? 0 : P.stage_per;
  const int G = gridDim.y;
  // this group's share of a unit's column values -> dead rows of the group's first column block
  auto stage = [&](int unit) {
    const cplx* src = zt + ((size_t)unit * P.np2 + z) * nvec + gid * per;
    const int cnt = min(per, nvec - gid * per);
    cplx* dst = pl + P.ksplit * pitch + gid * QB200_BLOCK_LINES;
    for (int j = g.tid; j < cnt; j += g.nthr) cp_async16(dst + (j >> 3) * pitch + (j & 7), src + j);
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
    __syncthreads();   // tables / staged values visible; the previous unit's readers are done with the plane
    if (OP != OP_FWD) {
      for (int i = threadIdx.x; i < nlo; i += blockDim.x) pl[i] = make_double2(0.0, 0.0);
      for (int i = threadIdx.x; i < nhi; i += blockDim.x) pl[hi0 + i] = make_double2(0.0, 0.0);
      __syncthreads();
      if (per) {
        const FastDiv dp(per);
        for (int i = threadIdx.x; i < nvec; i += blockDim.x) {
          int j;
          const int sh = dp.div(i, j);
          pl[colpos[i]] = pl[(P.ksplit + (j >> 3)) * pitch + sh * QB200_BLOCK_LINES + (j & 7)];
        }
      } else {
        constexpr int U = 8;                     // loads in flight per thread
        for (int i0 = threadIdx.x; i0 < nvec; i0 += U * blockDim.x) {
          cplx val[U];
#pragma unroll
          for (int u = 0; u < U; u++) val[u] = ztrow[min(i0 + u * (int)blockDim.x, nvec - 1)];
#pragma unroll
          for (int u = 0; u < U; u++) { const int i = i0 + u * blockDim.x; if (i < nvec) pl[colpos[i]] = val[u]; }
        }
      }
      __syncthreads();
      // x direction, whole CTA: the kept rows, digit-reversed -> natural
      fft_block_dit<+1>(cta, pl, P.nkeep, P.nkeep, rows, 1, P.f0, tw0, P.f0.nf - 1, false, all);
      __syncthreads();
    }
    // y direction: blocks of 8 columns, one group each
    for (int b = gid; b * QB200_BLOCK_LINES < np0; b += ngrp) {
      const int c0 = b * QB200_BLOCK_LINES;
      const int nc = min(QB200_BLOCK_LINES, np0 - c0);
      cplx* blk = pl + c0;
      const int nf = P.f1.nf;
      if (OP != OP_FWD && nf > 1) {
        fft_block_dif<+1>(g, blk, nc, QB200_BLOCK_LINES, colsmap, pitch, P.f1, tw1, 0, nf - 1, true, keepy);
        g.sync();
      }
      MidArgs a;
      a.v = v + (size_t)z * np01 + c0;
      a.rho = rho_part + (size_t)blockIdx.y * N + (size_t)z * np01 + c0;
      a.f = f + (size_t)unit * N + (size_t)z * np01 + c0;
      a.facu = facu; a.facv = facv; a.np0 = np0; a.zero_imag = zero_imag; a.exp = 0;
      mid_pass_any<OP>(rl, g, blk, nc, pitch, np1, yrev, a, nf == 1, keepy);
      if ((OP == OP_HPSI || OP == OP_FWD) && nf > 1) {
        g.sync();
        fft_block_dit<-1>(g, blk, nc, QB200_BLOCK_LINES, colsmap, pitch, P.f1, tw1, nf - 2, true, keepy);
      }
      if (per && b == gid && nxt < nunits) {
        g.sync();            // the whole group is done with the block before its dead rows are overwritten
        stage(nxt);
      }
    }
    if (OP == OP_HPSI || OP == OP_FWD) {
      __syncthreads();
      fft_block_dif<-1>(cta, pl, P.nkeep, P.nkeep, rows, 1, P.f0, tw0, 0, P.f0.nf, false, all);
      __syncthreads();
      for (int i = threadIdx.x; i < nvec; i += blockDim.x) ztrow[i] = pl[colpos[i]];
    }
    unit = nxt;
  }
}

}  // namespace qb200
