// qball_b200/csrc/zcol_kernels.cuh
// Second-generation z-column kernels: the sphere <-> column-form ends of the transform
// (vector_to_zvec / doublevector_to_zvec + the z FFTs of bwd, FourierTransform.cc:1624-1720, 584-700;
//  the z FFTs of fwd + 1/N scale + zvec_to_vector / zvec_to_doublevector, :1300-1361, 1666-1752), fused with
//  cp += ... and the kinetic term (SlaterDet.cc:1005-1036, EnergyFunctional.cc:1675-1690) on the way out.
//
// The first-generation kernels (transform_kernels.cuh) were latency bound: 16 resident warps per SM issuing
// dependent LDG -> STS pairs reached 1.4-1.5 TB/s (ncu: long_scoreboard 5-13 cycles per issue).  Here
//   * a CTA is PERSISTENT over units (states) for one fixed block of rods, grid (rod blocks, G) sized to one wave, so
//     the per-block tables (shared-memory position of every coefficient, 0.5|k+G|^2) are built once per CTA;
//   * every bulk global read is an asynchronous 16-byte copy (cp.async / LDGSTS) issued ONE UNIT AHEAD into a second
//     shared-memory buffer: the next unit's coefficients (bwd) or column tile (fwd) stream in while the current unit
//     is transformed, so there are always tens of KB in flight per SM and no load latency on the critical path;
//   * the column tile is kept z-major, lines[q][column] with an odd pitch: global reads/writes of zt[unit][z][iv] are
//     runs of `ncol` x 16 bytes, the line-fastest FFT tasks and the scatter/gather by digit-reversed z are
//     bank-conflict free;
//   * transforms use the group engine (fft_group.cuh) without un-permute passes: bwd scatters to digit-reversed z
//     and runs the DIT network (natural order out), fwd runs the DIF network and gathers from digit-reversed z.
#pragma once
#include "qb200_internal.h"
#include "fft_group.cuh"
#include "plane_static.cuh"

namespace qb200 {

// ZS selects the transform engine: DynZ = run-time shapes (group engine); ZShape<np2, columns per tile, zsplit, zskip> =
// compiled shape (pass_s of plane_static.cuh): the sphere only reaches |l| < zsplit, so the first backward pass knows
// which inputs are zero and the last forward pass which outputs are never gathered.
struct DynZ { static constexpr bool STATIC = false; };
template <int NP2_, int CB_, int ZSPLIT_, int ZSKIP_> struct ZShape {
  static constexpr bool STATIC = true;
  static constexpr int NP2 = NP2_, CB = CB_, PITCH = CB_ | 1, ZSPLIT = ZSPLIT_, ZSKIP = ZSKIP_;
};

__device__ __forceinline__ void zc_cp16(void* smem, const void* gmem)
{
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gmem) : "memory");
}
__device__ __forceinline__ void zc_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void zc_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

struct ZBlock {
  int col0, ncol, ig0, cnt;
};
__device__ __forceinline__ ZBlock zblock(const DevPlan& P, int rb)
{
  const int r0 = blockIdx.x * rb, r1 = min(r0 + rb, P.nrods);
  ZBlock b;
  b.col0 = P.is_real ? (r0 == 0 ? 0 : 2 * r0 - 1) : r0;
  b.ncol = (P.is_real ? 2 * r1 - 1 : r1) - b.col0;
  b.ig0 = P.rod_first[r0];
  b.cnt = (r1 < P.nrods ? P.rod_first[r1] : P.ngw) - b.ig0;
  return b;
}
// zq entry: digit-reversed z position (low 12 bits) and column (high bits) of a coefficient -> position in the tile
__device__ __forceinline__ int ztile_pos(int m, int pitch, int col0) { return (m & 4095) * pitch + (m >> 12) - col0; }

// ------------------------------------------------------------------------------------------------ backward
// grid (ceil(nrods/rb), G), block 256.  smem: tw[f2.twsize] | lines[np2*pitch] | stage[2][cper*cmax] | pos[cmax] (| posm[cmax])
template <int MODE, class ZS>
__global__ void __launch_bounds__(256, 2) k_zcol_bwd2(const __grid_constant__ DevPlan P, const cplx* __restrict__ c, size_t ldc,
                                                      cplx* __restrict__ zt, int nunits)
{
  extern __shared__ __align__(16) unsigned char smraw[];
  constexpr int CPER = MODE == MODE_PAIR ? 2 : 1;
  const int np2 = P.np2, CB = P.zb_cb, pitch = CB | 1, cmax = P.zb_cmax;
  cplx* tw = reinterpret_cast<cplx*>(smraw);
  cplx* lines = tw + P.f2.twsize;
  cplx* stage = lines + (size_t)np2 * pitch;
  int* pos = reinterpret_cast<int*>(stage + 2 * CPER * (size_t)cmax);
  int* posm = pos + cmax;
  const ZBlock b = zblock(P, P.zb_rb);
  const int tid = threadIdx.x, nthr = blockDim.x, G = gridDim.y;
  auto issue = [&](int unit, int buf) {
    const cplx* src = c + (size_t)unit * CPER * ldc + b.ig0;
    cplx* dst = stage + (size_t)buf * CPER * cmax;
    for (int e = tid; e < b.cnt; e += nthr) {
      zc_cp16(dst + e, src + e);
      if (MODE == MODE_PAIR) zc_cp16(dst + cmax + e, src + ldc + e);
    }
    zc_commit();
  };
  int unit = blockIdx.y;
  if (unit < nunits) issue(unit, 0);
  for (int i = tid; i < P.f2.twsize; i += nthr) tw[i] = P.tw2p[i];
  for (int e = tid; e < b.cnt; e += nthr) {
    pos[e] = ztile_pos(P.zq[b.ig0 + e], pitch, b.col0);
    if (P.is_real) posm[e] = ztile_pos(P.zqm[b.ig0 + e], pitch, b.col0);
  }
  const Grp g = { tid, nthr, 0 };
  const LineMap lm = { 1, 1 << 30, 0 };
  const Keep nokeep = { 0, 0 };
  const FastDiv dcol(b.ncol);
  for (int buf = 0; unit < nunits; unit += G, buf ^= 1) {
    // (a) the previous unit's write-out has finished reading `lines` once everybody is past this barrier
    __syncthreads();
    for (int i = tid; i < np2 * pitch; i += nthr) lines[i] = make_double2(0.0, 0.0);
    zc_wait_all();
    __syncthreads();                               // stage[buf] has landed for everybody; lines are zero
    if (unit + G < nunits) issue(unit + G, buf ^ 1);
    const cplx* s1 = stage + (size_t)buf * CPER * cmax;
    for (int e = tid; e < b.cnt; e += nthr) {
      const cplx a = s1[e];
      cplx pv, mv;
      if (MODE == MODE_PAIR) {
        const cplx bb = s1[cmax + e];
        pv = make_double2(a.x - bb.y, a.y + bb.x);
        mv = make_double2(a.x + bb.y, bb.x - a.y);
      } else {
        pv = a;
        mv = make_double2(a.x, -a.y);
      }
      lines[pos[e]] = pv;
      if (P.is_real) lines[posm[e]] = mv;          // same thread, later: the conjugate wins at G=0, as in the reference
    }
    __syncthreads();
    if constexpr (ZS::STATIC) {
      constexpr FftDesc FZ = make_fft_desc(ZS::NP2);
      dit_s<+1, ZS::NP2, ZS::PITCH, ZS::CB, ColsOff, ZS::ZSPLIT, ZS::ZSKIP, true, false, FZ.nf - 1>(tid, nthr, lines, tw, [] { __syncthreads(); });
    } else {
      fft_block_dit<+1>(g, lines, b.ncol, CB, lm, pitch, P.f2, tw, P.f2.nf - 1, false, nokeep);
    }
    __syncthreads();
    cplx* out = zt + (size_t)unit * np2 * P.nvec + b.col0;
    for (int e = tid; e < b.ncol * np2; e += nthr) {
      int lc;
      const int z = dcol.div(e, lc);
      out[(size_t)z * P.nvec + lc] = lines[z * pitch + lc];
    }
  }
  zc_wait_all();
}

// ------------------------------------------------------------------------------------------------ forward
// grid (ceil(nrods/rb), G), block 256.  smem: tw | lines[2][np2*pitch] | kpg2h[cmax] | pos[cmax] (| posm[cmax])
template <int MODE, class ZS>
__global__ void __launch_bounds__(256, 2) k_zcol_fwd2(const __grid_constant__ DevPlan P, const cplx* __restrict__ zt, cplx* __restrict__ out,
                                                      size_t ldc, int accumulate, const double* __restrict__ kpg2,
                                                      const cplx* __restrict__ cin, double scale, int nunits)
{
  extern __shared__ __align__(16) unsigned char smraw[];
  constexpr int CPER = MODE == MODE_PAIR ? 2 : 1;
  const int np2 = P.np2, CB = P.zf_cb, pitch = CB | 1, cmax = P.zf_cmax;
  cplx* tw = reinterpret_cast<cplx*>(smraw);
  cplx* lines0 = tw + P.f2.twsize;
  double* kh = reinterpret_cast<double*>(lines0 + 2 * (size_t)np2 * pitch);
  int* pos = reinterpret_cast<int*>(kh + cmax);
  int* posm = pos + cmax;
  const ZBlock b = zblock(P, P.zf_rb);
  const int tid = threadIdx.x, nthr = blockDim.x, G = gridDim.y;
  const FastDiv dcol(b.ncol);
  auto issue = [&](int unit, int buf) {
    const cplx* src = zt + (size_t)unit * np2 * P.nvec + b.col0;
    cplx* dst = lines0 + (size_t)buf * np2 * pitch;
    for (int e = tid; e < b.ncol * np2; e += nthr) {
      int lc;
      const int z = dcol.div(e, lc);
      zc_cp16(dst + z * pitch + lc, src + (size_t)z * P.nvec + lc);
    }
    zc_commit();
  };
  int unit = blockIdx.y;
  if (unit < nunits) issue(unit, 0);
  for (int i = tid; i < P.f2.twsize; i += nthr) tw[i] = P.tw2p[i];
  for (int e = tid; e < b.cnt; e += nthr) {
    pos[e] = ztile_pos(P.zq[b.ig0 + e], pitch, b.col0);
    if (MODE == MODE_PAIR) posm[e] = ztile_pos(P.zqm[b.ig0 + e], pitch, b.col0);
    kh[e] = kpg2 ? 0.5 * kpg2[b.ig0 + e] : 0.0;
  }
  const Grp g = { tid, nthr, 0 };
  const LineMap lm = { 1, 1 << 30, 0 };
  const Keep nokeep = { 0, 0 };
  const double hs = 0.5 * scale;
  for (int buf = 0; unit < nunits; unit += G, buf ^= 1) {
    zc_wait_all();
    __syncthreads();          // this unit's tile has landed; everybody has finished the previous unit's gather (other buffer)
    if (unit + G < nunits) issue(unit + G, buf ^ 1);
    cplx* lines = lines0 + (size_t)buf * np2 * pitch;
    if constexpr (ZS::STATIC) {
      constexpr FftDesc FZ = make_fft_desc(ZS::NP2);
      dif_s<-1, ZS::NP2, ZS::PITCH, ZS::CB, ColsOff, ZS::ZSPLIT, ZS::ZSKIP, false, true, 0, FZ.nf - 1>(tid, nthr, lines, tw, [] { __syncthreads(); });
    } else {
      fft_block_dif<-1>(g, lines, b.ncol, CB, lm, pitch, P.f2, tw, 0, P.f2.nf, false, nokeep);
    }
    __syncthreads();
    const size_t s1 = (size_t)unit * CPER * ldc + b.ig0;
    cplx* o1 = out + s1;
    const cplx* i1 = cin + s1;
    constexpr int U = MODE == MODE_PAIR ? 2 : 4;
    for (int e0 = tid; e0 < b.cnt; e0 += U * nthr) {
      cplx a1[U], a2[U], q1[U], q2[U];
#pragma unroll
      for (int u = 0; u < U; u++) {
        const int e = e0 + u * nthr;
        if (e < b.cnt) {
          if (kpg2) { a1[u] = i1[e]; if (MODE == MODE_PAIR) a2[u] = i1[ldc + e]; }
          if (accumulate) { q1[u] = o1[e]; if (MODE == MODE_PAIR) q2[u] = o1[ldc + e]; }
        }
      }
#pragma unroll
      for (int u = 0; u < U; u++) {
        const int e = e0 + u * nthr;
        if (e < b.cnt) {
          const cplx pv = lines[pos[e]];
          cplx w1, w2;
          if (MODE == MODE_PAIR) {
            const cplx mv = lines[posm[e]];
            w1 = make_double2(hs * (pv.x + mv.x), hs * (pv.y - mv.y));
            w2 = make_double2(hs * (pv.y + mv.y), hs * (mv.x - pv.x));
          } else {
            w1 = make_double2(scale * pv.x, scale * pv.y);
          }
          if (kpg2) {
            const double h = kh[e];
            w1.x += h * a1[u].x; w1.y += h * a1[u].y;
            if (MODE == MODE_PAIR) { w2.x += h * a2[u].x; w2.y += h * a2[u].y; }
          }
          if (accumulate) {
            w1.x += q1[u].x; w1.y += q1[u].y;
            if (MODE == MODE_PAIR) { w2.x += q2[u].x; w2.y += q2[u].y; }
          }
          o1[e] = w1;
          if (MODE == MODE_PAIR) o1[ldc + e] = w2;
        }
      }
    }
  }
  zc_wait_all();
}

}  // namespace qb200
