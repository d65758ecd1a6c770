// qball_b200/csrc/nonlocal.cu -- NonLocalPotential::energy, norm-conserving branch
// (/root/reference/src/qball/NonLocalPotential.cc:1909-2171, 2628-2643) as two fused FP64 tensor-core GEMMs.
//
// The reference materialises anl[ig,(ia,ipr)] = twnl[ipr][ig] * (-i)^l * exp(-i (k+G).tau_ia)  (:1959-2036) for blocks
// of <=128 atoms and calls BLAS:  fnl = anl^H c (:2050-2068),  E_nl += occ wt/omega |fnl|^2 (:2106-2148),
// cp += anl (wt/omega fnl) (:2150-2171).  Here anl is never stored: both GEMM kernels regenerate their anl tiles on the
// fly (one FP64 sincos per (atom, G) per tile, shared by the atom's projectors) straight into shared memory and feed
// mma.sync.m8n8k4.f64 (DMMA -- tcgen05 has no FP64 kind).  Complex arithmetic is mapped on real DMMA:
//   k_fnl : out[(p,re|im), n] = sum_k A[(p,.)][k] B[k][n],  k over the 2*ngw reals (g,re),(g,im), B = c as stored
//   k_back: cp[(g,re|im), n] += sum_{(p,re|im)} A2[(g,.)][(p,.)] f'[(p,.), n]
// At the Gamma point both are plain real GEMMs over the 2*ngw reals with the G=0 half weight (:2070-2082) folded
// into k_fnl's tile generation and the factor 2 (:2102) into the epilogue.
// Internal projector order is atom-major, p = ia*npr + ipr (the reference's ia + ipr*nab is never exposed).
// Summation is deterministic: split-K partials are reduced in fixed order, E_nl by a fixed tree.
#include "qb200_internal.h"
#include <algorithm>
#include <cmath>
#include <cstdio>

namespace qb200 {

// Tile geometry of both GEMM kernels: CTA = 512 threads = 16 warps (4 x 4), warp tile 32 x 32 (16 m8n8 accumulators),
// CTA tile 128 x 128 reals, 32 reals of the reduction dimension per stage, two stages of shared memory:
// while the tensor pipe works on stage s, the FP64 pipe generates the anl tile of stage s+1 (sincos) and cp.async
// brings in its B tile -- one __syncthreads per stage.
#define NL_TM 128
#define NL_TN 128
#define NL_KSTEP 32
#define NL_THREADS 512
#define NL_PITCH 36            // doubles per row of a [row][k] tile: 32 + 4 -> conflict-free DMMA fragment loads
#define NL_PITCH_KR 132        // doubles per k-row of a [k][row] tile: 128 + 4 (same property, k-major)
#define NL_NPRMAX 32           // projectors per atom whose twnl values are staged in shared memory (more: read from global)
#define NL_TAUMAX 1024         // atoms whose positions are staged in shared memory by k_back (more: read from global)

__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b)
{
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
// 16/8-byte asynchronous copies global -> shared; !valid copies nothing and writes zeros
__device__ __forceinline__ void nl_cp16(void* smem, const void* gmem, bool valid)
{
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  const int n = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(sa), "l"(gmem), "r"(n) : "memory");
}
__device__ __forceinline__ void nl_cp8(void* smem, const void* gmem, bool valid)
{
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  const int n = valid ? 8 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(sa), "l"(gmem), "r"(n) : "memory");
}
__device__ __forceinline__ void nl_cp4(void* smem, const void* gmem, bool valid)
{
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  const int n = valid ? 4 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(sa), "l"(gmem), "r"(n) : "memory");
}
__device__ __forceinline__ void nl_cp_wait() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// anl = t * (-i)^l * (c + i s)     (NonLocalPotential.cc:2002-2034)
__device__ __forceinline__ double2 anl_value(int l, double t, double s, double c)
{
  switch (l & 3) {
    case 0: return make_double2(t * c, t * s);
    case 1: return make_double2(t * s, -t * c);
    case 2: return make_double2(-t * c, -t * s);
    default: return make_double2(-t * s, t * c);
  }
}

struct NlSpecies {
  int na, npr, M;                 // M = na*npr
  const int* lproj;               // [npr]
  const double* wt;               // [npr]
  const double* twnl;             // [npr][ngw]
  const double* tau;              // [na][3]
  const double2* ph;              // [na][JT] separable phase tables (NlLattice), or null
};

// Optional integer description of the plane waves (qb200_nl_set_lattice): k+G = kpoint + h b0 + k b1 + l b2, so
//   exp(-i (k+G).tau) = [exp(-i kpoint.tau) exp(-i h b0.tau)] * exp(-i k b1.tau) * exp(-i l b2.tau)
// and the FP64 sincos per (atom, G) of the tile generation (which runs on the same FP64 units as DMMA) becomes three
// table look-ups and two complex multiplications.  Tables: per atom JT = J0+J1+J2 entries, Jd = 2*jmax[d]+1.
struct NlLattice {
  const int* idx;                 // [3][ngw] (h, k, l planes), null: sincos path
  int jmax[3];
  int JT;
};

__device__ __forceinline__ double2 nl_phase(const NlLattice& L, const double2* __restrict__ T, int h, int k, int l)
{
  const double2 a = __ldg(T + h + L.jmax[0]);
  const double2 b = __ldg(T + 2 * L.jmax[0] + 1 + k + L.jmax[1]);
  const double2 c = __ldg(T + 2 * L.jmax[0] + 1 + 2 * L.jmax[1] + 1 + l + L.jmax[2]);
  const double2 ab = make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
  return make_double2(ab.x * c.x - ab.y * c.y, ab.x * c.y + ab.y * c.x);
}

// grid (na), block 128: tables of one species.  bt[d] = b_d . tau, kt = kpoint . tau
__global__ void k_phase_tables(NlSpecies S, NlLattice L, double b00, double b01, double b02, double b10, double b11, double b12,
                               double b20, double b21, double b22, double k0, double k1, double k2, double2* __restrict__ out)
{
  const int ia = blockIdx.x;
  const double tx = S.tau[3 * ia], ty = S.tau[3 * ia + 1], tz = S.tau[3 * ia + 2];
  const double bt[3] = { b00 * tx + b01 * ty + b02 * tz, b10 * tx + b11 * ty + b12 * tz, b20 * tx + b21 * ty + b22 * tz };
  const double kt = k0 * tx + k1 * ty + k2 * tz;
  double2* T = out + (size_t)ia * L.JT;
  int off = 0;
  for (int d = 0; d < 3; d++) {
    const int J = 2 * L.jmax[d] + 1;
    for (int j = threadIdx.x; j < J; j += blockDim.x) {
      const double arg = -((double)(j - L.jmax[d]) * bt[d] + (d == 0 ? kt : 0.0));
      double sn, cs;
      sincos(arg, &sn, &cs);
      T[off + j] = make_double2(cs, sn);
    }
    off += J;
  }
}

// one stage of the warp tile: acc[i][j] += A(32 x 32 reals) * B(32 x 32 reals)
template <bool A_KMAJOR>
__device__ __forceinline__ void warp_mma_stage(const double* __restrict__ As, const double* __restrict__ Bs, double (&acc)[4][4][2],
                                               int lane, int wm, int wn)
{
  const int r = lane >> 2, kq = lane & 3;
  const double* a0 = A_KMAJOR ? As + kq * NL_PITCH_KR + wm * 32 + r : As + (wm * 32 + r) * NL_PITCH + kq;
  const double* b0 = Bs + (wn * 32 + r) * NL_PITCH + kq;
#pragma unroll
  for (int k4 = 0; k4 < NL_KSTEP / 4; k4++) {
    double a[4], b[4];
#pragma unroll
    for (int i = 0; i < 4; i++) a[i] = A_KMAJOR ? a0[k4 * 4 * NL_PITCH_KR + i * 8] : a0[i * 8 * NL_PITCH + k4 * 4];
#pragma unroll
    for (int j = 0; j < 4; j++) b[j] = b0[j * 8 * NL_PITCH + k4 * 4];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
      for (int j = 0; j < 4; j++) dmma(acc[i][j][0], acc[i][j][1], a[i], b[j]);
  }
}

// shared-memory layout of k_fnl (doubles)
#define FNL_AS 0
#define FNL_BS (FNL_AS + 2 * NL_TM * NL_PITCH)
#define FNL_STG (FNL_BS + 2 * NL_TN * NL_PITCH)                 // 2 x [(3 + NL_NPRMAX)][16]: kpgx and twnl of a stage
#define FNL_TAU (FNL_STG + 2 * (3 + NL_NPRMAX) * 16)            // [128][3]
#define FNL_IDX (FNL_TAU + 3 * 128)                             // 2 x [3][16] ints: (h,k,l) of a stage's plane waves
#define FNL_LPR (FNL_IDX + 2 * 3 * 16 / 2)                      // [NL_NPRMAX] ints: lproj
#define FNL_SMEM_BYTES ((FNL_LPR + NL_NPRMAX / 2) * 8)

// ------------------------------------------------------------------------------------------------ fnl = anl^H c
// grid (ceil(M/PT), ceil(nst/128), ksplit), PT = 64 projectors (complex: rows (p,re),(p,im)) or 128 (Gamma: real).
// part[(ks*ncols + col)*Mp + p], ncols = IS_REAL ? nst : 2*nst, col = n or 2n+{re,im}.
// Reduction over the reals (g,re),(g,im) of the plane waves of this CTA's chunk, 16 plane waves per stage:
//   A[(p,re)][(g,.)] = ( a.x, a.y),  A[(p,im)][(g,.)] = (-a.y, a.x),  B[(g,.)][n] = c[g,n]   (a = anl[g,p]; fnl = conj(a) c)
template <int IS_REAL>
__global__ void __launch_bounds__(NL_THREADS, 1) k_fnl(NlSpecies S, NlLattice L, int ngw, const double* __restrict__ kpgx,
                                                         const double2* __restrict__ c, size_t ldc, int nst, int gchunk,
                                                         double* __restrict__ part, int Mp)
{
  extern __shared__ __align__(16) double nl_smem[];
  double* As = nl_smem + FNL_AS;
  double* Bs = nl_smem + FNL_BS;
  double* stg = nl_smem + FNL_STG;
  double* taus = nl_smem + FNL_TAU;
  int* istg = reinterpret_cast<int*>(nl_smem + FNL_IDX);
  int* lprs = reinterpret_cast<int*>(nl_smem + FNL_LPR);
  const bool tables = L.idx != nullptr && S.ph != nullptr;
  constexpr int PT = IS_REAL ? NL_TM : NL_TM / 2;
  constexpr int STG = (3 + NL_NPRMAX) * 16;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp >> 2, wn = warp & 3;
  const int p0 = blockIdx.x * PT, n0 = blockIdx.y * NL_TN;
  const int ncols = IS_REAL ? nst : 2 * nst;
  const int g0 = blockIdx.z * gchunk, g1 = min(g0 + gchunk, ngw);
  const int npr = S.npr;
  const int ia0 = p0 / npr;
  const int ia1 = min((p0 + PT - 1) / npr, S.na - 1);
  const int nat = ia1 - ia0 + 1;
  const int nstage = (g1 - g0 + 15) / 16;
  const bool tw_staged = npr <= NL_NPRMAX;

  for (int i = tid; i < 2 * NL_TM * NL_PITCH; i += NL_THREADS) As[i] = 0.0;     // rows no projector maps to stay zero
  for (int i = tid; i < 3 * nat; i += NL_THREADS) taus[i] = S.tau[3 * ia0 + i];
  for (int i = tid; i < min(npr, NL_NPRMAX); i += NL_THREADS) lprs[i] = S.lproj[i];     // no global load inside the stage loop

  auto issue_stg = [&](int st) {                 // kpgx and twnl values of stage st -> stg[st & 1]
    if (st >= nstage) return;
    double* dst = stg + (st & 1) * STG;
    const int gs = g0 + st * 16;
    const int nrow = 3 + (tw_staged ? npr : 0);
    for (int i = tid; i < nrow * 16; i += NL_THREADS) {
      const int row = i >> 4, gl = i & 15, g = gs + gl;
      const bool ok = g < g1;
      const double* src = row < 3 ? kpgx + (size_t)row * ngw + (ok ? g : 0) : S.twnl + (size_t)(row - 3) * ngw + (ok ? g : 0);
      nl_cp8(dst + i, src, ok);
    }
    if (tables && tid >= NL_THREADS - 48) {
      const int i = tid - (NL_THREADS - 48), g = gs + (i & 15);
      const bool ok = g < g1;
      nl_cp4(istg + (st & 1) * 48 + i, L.idx + (size_t)(i >> 4) * ngw + (ok ? g : 0), ok);
    }
  };
  auto issue_b = [&](int st) {                   // c[g, n] of stage st -> Bs[st & 1][n][(g,re),(g,im)]
    if (st >= nstage) return;
    double* dst = Bs + (st & 1) * NL_TN * NL_PITCH;
    const int gs = g0 + st * 16;
    for (int i = tid; i < NL_TN * 16; i += NL_THREADS) {
      const int nl = i >> 4, gl = i & 15, n = n0 + nl, g = gs + gl;
      const bool ok = n < nst && g < g1;
      nl_cp16(dst + nl * NL_PITCH + 2 * gl, c + (ok ? (size_t)n * ldc + g : 0), ok);
    }
  };
  auto generate = [&](int st) {                  // anl tile of stage st -> As[st & 1]
    if (st >= nstage) return;
    double* A = As + (st & 1) * NL_TM * NL_PITCH;
    const double* sg = stg + (st & 1) * STG;
    const int gs = g0 + st * 16;
    for (int w = tid; w < nat * 16; w += NL_THREADS) {
      const int gl = w & 15, ai = w >> 4, g = gs + gl;
      const bool ok = g < g1;
      double sn = 0.0, cs = 0.0;
      if (ok) {
        if (tables) {
          const int* ig = istg + (st & 1) * 48;
          const double2 e = nl_phase(L, S.ph + (size_t)(ia0 + ai) * L.JT, ig[gl], ig[16 + gl], ig[32 + gl]);
          cs = e.x; sn = e.y;
        } else {
          const double arg = -(sg[gl] * taus[3 * ai] + sg[16 + gl] * taus[3 * ai + 1] + sg[32 + gl] * taus[3 * ai + 2]);
          sincos(arg, &sn, &cs);
        }
      }
      for (int ipr = 0; ipr < npr; ipr++) {
        const int pl = (ia0 + ai) * npr + ipr - p0;
        if (pl < 0 || pl >= PT || p0 + pl >= S.M) continue;
        double2 a = make_double2(0.0, 0.0);
        if (ok) {
          const double t = tw_staged ? sg[(3 + ipr) * 16 + gl] : S.twnl[(size_t)ipr * ngw + g];
          a = anl_value(tw_staged ? lprs[ipr] : S.lproj[ipr], t, sn, cs);
          if (IS_REAL && g == 0) a.x *= 0.5;       // G=0 counted once: dger fix, NonLocalPotential.cc:2078-2080
        }
        if (IS_REAL) {
          *reinterpret_cast<double2*>(A + pl * NL_PITCH + 2 * gl) = a;
        } else {
          *reinterpret_cast<double2*>(A + (2 * pl) * NL_PITCH + 2 * gl) = a;
          *reinterpret_cast<double2*>(A + (2 * pl + 1) * NL_PITCH + 2 * gl) = make_double2(-a.y, a.x);
        }
      }
    }
  };

  double acc[4][4][2];
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) acc[i][j][0] = acc[i][j][1] = 0.0;

  issue_stg(0);
  issue_b(0);
  nl_cp_wait();
  __syncthreads();
  issue_stg(1);
  generate(0);
  nl_cp_wait();
  __syncthreads();
  for (int st = 0; st < nstage; st++) {
    // generation first: before this stage's DMMAs (the warps that hold work items start multiplying late, the others
    // keep the FP64/DMMA pipe full meanwhile -- one warp per sub-partition saturates it) and before the cp.async
    // issue (otherwise the phase-table loads share a scoreboard with the copies and wait for HBM)
    generate(st + 1);
    issue_b(st + 1);
    issue_stg(st + 2);
    warp_mma_stage<false>(As + (st & 1) * NL_TM * NL_PITCH, Bs + (st & 1) * NL_TN * NL_PITCH, acc, lane, wm, wn);
    nl_cp_wait();
    __syncthreads();
  }
  const int r = lane >> 2, cq = lane & 3;
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++)
#pragma unroll
      for (int e = 0; e < 2; e++) {
        const int row = wm * 32 + i * 8 + r;
        const int n = n0 + wn * 32 + j * 8 + 2 * cq + e;
        const int p = p0 + (IS_REAL ? row : (row >> 1));
        const int col = IS_REAL ? n : 2 * n + (row & 1);
        if (p < S.M && n < nst) part[((size_t)blockIdx.z * ncols + col) * Mp + p] = acc[i][j][e];
      }
}

// ------------------------------------------------------------------------------------------------ E_nl, fnl <- wt/omega fnl
// one thread per (n, p); fs[n][p] complex (or real at Gamma); block partial sums of E_nl to eblk
template <int IS_REAL>
__global__ void __launch_bounds__(256) k_fnl_finish(NlSpecies S, const double* __restrict__ part, int Mp, int nst, int ksplit,
                                                    const double* __restrict__ occ, double omega_inv,
                                                    double* __restrict__ fs, double* __restrict__ eblk)
{
  __shared__ double red[256];
  const size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  const size_t total = (size_t)nst * S.M;
  double e = 0.0;
  if (idx < total) {
    const int n = (int)(idx / S.M), p = (int)(idx % S.M);
    const int ncols = IS_REAL ? nst : 2 * nst;
    const double fac = S.wt[p % S.npr] * omega_inv;
    if (IS_REAL) {
      double f = 0.0;
      for (int ks = 0; ks < ksplit; ks++) f += part[((size_t)ks * ncols + n) * Mp + p];
      f *= 2.0;                                           // G and -G, NonLocalPotential.cc:2102
      e = fac * occ[n] * f * f;
      fs[(size_t)n * Mp + p] = fac * f;
    } else {
      double fr = 0.0, fi = 0.0;
      for (int ks = 0; ks < ksplit; ks++) {
        fr += part[((size_t)ks * ncols + 2 * n) * Mp + p];
        fi += part[((size_t)ks * ncols + 2 * n + 1) * Mp + p];
      }
      e = fac * occ[n] * (fr * fr + fi * fi);
      fs[2 * ((size_t)n * Mp + p)] = fac * fr;
      fs[2 * ((size_t)n * Mp + p) + 1] = fac * fi;
    }
  }
  red[threadIdx.x] = e;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) eblk[blockIdx.x] = red[0];
}

__global__ void k_sum_blocks(const double* __restrict__ eblk, int n, double* __restrict__ acc)
{
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    double s = 0.0;
    for (int i = 0; i < n; i++) s += eblk[i];
    *acc += s;
  }
}

// shared-memory layout of k_back (doubles)
#define BK_AS 0
#define BK_BS (BK_AS + 2 * NL_KSTEP * NL_PITCH_KR)
#define BK_KP (BK_BS + 2 * NL_TN * NL_PITCH)                    // [3][64] kpgx of the CTA's plane waves
#define BK_TW (BK_KP + 3 * 64)                                  // [NL_NPRMAX][64] twnl of the CTA's plane waves
#define BK_TAU (BK_TW + NL_NPRMAX * 64)                         // [NL_TAUMAX][3]
#define BK_IDX (BK_TAU + 3 * NL_TAUMAX)                         // [3][64] ints: (h,k,l) of the CTA's plane waves
#define BK_LPR (BK_IDX + 3 * 64 / 2)                            // [NL_NPRMAX] ints: lproj
#define BK_SMEM_BYTES ((BK_LPR + NL_NPRMAX / 2) * 8)

// ------------------------------------------------------------------------------------------------ cp += anl * fs
// grid (ceil(ngw/64), ceil(nst/128)): 128 output reals (64 plane waves x re/im) x 128 states per CTA.
// Reduction over the projectors, PSTEP = 16 complex (Gamma: 32 real) per stage; A is stored k-major ([k][row]) so that
// the generating thread of (atom, g) writes (re,im) row pairs with one 16-byte store, conflict-free:
//   complex: A[(g,re)][(p,re)] = a.x  A[(g,re)][(p,im)] = -a.y  A[(g,im)][(p,re)] = a.y  A[(g,im)][(p,im)] = a.x
//   Gamma:   A[(g,re)][p] = a.x  A[(g,im)][p] = a.y            B[k][n] = fs[n][k]  (fs = wt/omega * fnl)
template <int IS_REAL>
__global__ void __launch_bounds__(NL_THREADS, 1) k_back(NlSpecies S, NlLattice L, int ngw, const double* __restrict__ kpgx,
                                                          const double* __restrict__ fs, int Mp, double2* __restrict__ cp,
                                                          size_t ldc, int nst)
{
  extern __shared__ __align__(16) double nl_smem[];
  double* As = nl_smem + BK_AS;
  double* Bs = nl_smem + BK_BS;
  double* kps = nl_smem + BK_KP;
  double* tws = nl_smem + BK_TW;
  double* taus = nl_smem + BK_TAU;
  int* idxs = reinterpret_cast<int*>(nl_smem + BK_IDX);
  int* lprs = reinterpret_cast<int*>(nl_smem + BK_LPR);
  const bool tables = L.idx != nullptr && S.ph != nullptr;
  constexpr int PSTEP = IS_REAL ? NL_KSTEP : NL_KSTEP / 2;   // projectors per stage
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp >> 2, wn = warp & 3;
  const int g0 = blockIdx.x * 64, n0 = blockIdx.y * NL_TN;
  const int npr = S.npr;
  const int nstage = (S.M + PSTEP - 1) / PSTEP;
  const bool tw_staged = npr <= NL_NPRMAX, tau_staged = S.na <= NL_TAUMAX;

  for (int i = tid; i < 3 * 64; i += NL_THREADS) { const int g = g0 + (i & 63); kps[i] = g < ngw ? kpgx[(size_t)(i >> 6) * ngw + g] : 0.0; }
  if (tw_staged)
    for (int i = tid; i < npr * 64; i += NL_THREADS) { const int g = g0 + (i & 63); tws[i] = g < ngw ? S.twnl[(size_t)(i >> 6) * ngw + g] : 0.0; }
  for (int i = tid; i < min(npr, NL_NPRMAX); i += NL_THREADS) lprs[i] = S.lproj[i];
  if (tau_staged && !tables) for (int i = tid; i < 3 * S.na; i += NL_THREADS) taus[i] = S.tau[i];
  if (tables) for (int i = tid; i < 3 * 64; i += NL_THREADS) { const int g = g0 + (i & 63); idxs[i] = g < ngw ? L.idx[(size_t)(i >> 6) * ngw + g] : 0; }
  for (int i = tid; i < 2 * NL_KSTEP * NL_PITCH_KR; i += NL_THREADS) As[i] = 0.0;

  auto issue_b = [&](int st) {                   // fs of stage st -> Bs[st & 1][n][k]
    if (st >= nstage) return;
    double* dst = Bs + (st & 1) * NL_TN * NL_PITCH;
    const int ps = st * PSTEP;
    for (int i = tid; i < NL_TN * 16; i += NL_THREADS) {
      const int nl = i >> 4, ch = i & 15, n = n0 + nl;          // chunk ch = doubles 2ch, 2ch+1 of the stage's 32
      bool ok;
      const double* src;
      if (IS_REAL) { ok = n < nst && ps + 2 * ch < Mp; src = fs + (ok ? (size_t)n * Mp + ps + 2 * ch : 0); }
      else { ok = n < nst && ps + ch < S.M; src = fs + (ok ? 2 * ((size_t)n * Mp + ps + ch) : 0); }
      nl_cp16(dst + nl * NL_PITCH + 2 * ch, src, ok);
    }
  };
  auto generate = [&](int st) {                  // anl tile of stage st -> As[st & 1], k-major
    if (st >= nstage) return;
    double* A = As + (st & 1) * NL_KSTEP * NL_PITCH_KR;
    const int ps = st * PSTEP;
    const int ia0 = ps / npr;
    const int ia1 = min((ps + PSTEP - 1) / npr, S.na - 1);
    const int nat = ia1 - ia0 + 1;
    if (ps + PSTEP > S.M) {                      // last stage: projector slots past M must read as zero
      const int k0 = (IS_REAL ? 1 : 2) * (S.M - ps);
      for (int i = tid; i < (NL_KSTEP - k0) * NL_PITCH_KR; i += NL_THREADS) A[k0 * NL_PITCH_KR + i] = 0.0;
    }
    for (int w = tid; w < nat * 64; w += NL_THREADS) {
      const int gl = w & 63, ia = ia0 + (w >> 6), g = g0 + gl;
      const bool ok = g < ngw;
      double sn = 0.0, cs = 0.0;
      if (ok) {
        if (tables) {
          const double2 e = nl_phase(L, S.ph + (size_t)ia * L.JT, idxs[gl], idxs[64 + gl], idxs[128 + gl]);
          cs = e.x; sn = e.y;
        } else {
          const double* t3 = tau_staged ? taus + 3 * ia : S.tau + 3 * ia;
          const double arg = -(kps[gl] * t3[0] + kps[64 + gl] * t3[1] + kps[128 + gl] * t3[2]);
          sincos(arg, &sn, &cs);
        }
      }
      for (int ipr = 0; ipr < npr; ipr++) {
        const int pl = ia * npr + ipr - ps;
        if (pl < 0 || pl >= PSTEP || ps + pl >= S.M) continue;
        double2 a = make_double2(0.0, 0.0);
        if (ok) {
          const double t = tw_staged ? tws[ipr * 64 + gl] : S.twnl[(size_t)ipr * ngw + g];
          a = anl_value(tw_staged ? lprs[ipr] : S.lproj[ipr], t, sn, cs);
        }
        if (IS_REAL) {
          *reinterpret_cast<double2*>(A + pl * NL_PITCH_KR + 2 * gl) = a;
        } else {
          *reinterpret_cast<double2*>(A + (2 * pl) * NL_PITCH_KR + 2 * gl) = a;
          *reinterpret_cast<double2*>(A + (2 * pl + 1) * NL_PITCH_KR + 2 * gl) = make_double2(-a.y, a.x);
        }
      }
    }
  };

  double acc[4][4][2];
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) acc[i][j][0] = acc[i][j][1] = 0.0;

  issue_b(0);
  __syncthreads();            // tables and the zeroed A buffers are visible
  generate(0);
  nl_cp_wait();
  __syncthreads();
  for (int st = 0; st < nstage; st++) {
    generate(st + 1);
    issue_b(st + 1);
    warp_mma_stage<true>(As + (st & 1) * NL_KSTEP * NL_PITCH_KR, Bs + (st & 1) * NL_TN * NL_PITCH, acc, lane, wm, wn);
    nl_cp_wait();
    __syncthreads();
  }
  const int r = lane >> 2, cq = lane & 3;
  double* cpd = reinterpret_cast<double*>(cp);
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++)
#pragma unroll
      for (int e = 0; e < 2; e++) {
        const int row = wm * 32 + i * 8 + r;                 // real row within the tile: 2*gl + (re|im)
        const int g = g0 + (row >> 1);
        const int n = n0 + wn * 32 + j * 8 + 2 * cq + e;
        if (g < ngw && n < nst) cpd[2 * ((size_t)n * ldc + g) + (row & 1)] += acc[i][j][e];
      }
}

}  // namespace qb200

using namespace qb200;

struct qb200_nl {
  int device;
  cudaStream_t stream;
  int ngw, is_real;
  double omega;
  double* kpgx;                                // device [3][ngw]
  std::vector<NlSpecies> sp;
  std::vector<void*> owned;
  double *part, *fs, *eblk, *occ_dev, *enl_dev; size_t part_cap, fs_cap, eblk_cap, occ_cap;
  double *st_c, *st_cp; size_t st_c_cap, st_cp_cap;
  long long launches;
  int nsm;
  // optional lattice description (qb200_nl_set_lattice)
  NlLattice lat;
  double bvec[9], kcart[3];
  std::vector<double2*> ph;                    // per species phase tables
  bool ph_dirty;
};

static int nl_ensure(double** buf, size_t* cap, size_t elems)
{
  if (*cap >= elems && *buf) return QB200_OK;
  if (*buf) { cudaFree(*buf); *buf = nullptr; *cap = 0; }
  QB_CUDA(cudaMalloc((void**)buf, std::max<size_t>(elems, 1) * sizeof(double)));
  *cap = elems;
  return QB200_OK;
}

template <class T> static int nl_upload(qb200_nl* nl, const T* h, size_t n, const T** d)
{
  void* p = nullptr;
  QB_CUDA(cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(T)));
  nl->owned.push_back(p);
  if (n) QB_CUDA(cudaMemcpy(p, h, n * sizeof(T), cudaMemcpyHostToDevice));
  *d = (const T*)p;
  return QB200_OK;
}

extern "C" int qb200_nl_create(qb200_nl** out, int device, int ngw, int is_real, double omega, const double* kpgx)
{
  if (!out || ngw < 1 || !(omega > 0.0) || !kpgx) { set_error("qb200_nl_create: bad argument"); return QB200_EINVAL; }
  *out = nullptr;
  int ndev = 0;
  QB_CUDA(cudaGetDeviceCount(&ndev));
  if (device < 0 || device >= ndev) { set_error("qb200_nl_create: no such CUDA device"); return QB200_ENODEV; }
  QB_CUDA(cudaSetDevice(device));
  qb200_nl* nl = new qb200_nl();
  nl->device = device; nl->stream = 0; nl->ngw = ngw; nl->is_real = is_real ? 1 : 0; nl->omega = omega;
  nl->part = nl->fs = nl->eblk = nl->occ_dev = nl->enl_dev = nl->st_c = nl->st_cp = nullptr;
  nl->part_cap = nl->fs_cap = nl->eblk_cap = nl->occ_cap = nl->st_c_cap = nl->st_cp_cap = 0;
  nl->launches = 0;
  nl->lat.idx = nullptr; nl->lat.jmax[0] = nl->lat.jmax[1] = nl->lat.jmax[2] = 0; nl->lat.JT = 0;
  nl->ph_dirty = true;
  cudaDeviceProp prop;
  QB_CUDA(cudaGetDeviceProperties(&prop, device));
  nl->nsm = prop.multiProcessorCount;
  const double* d;
  int rc = nl_upload(nl, kpgx, 3 * (size_t)ngw, &d);
  if (rc) { qb200_nl_destroy(nl); return rc; }
  nl->kpgx = const_cast<double*>(d);
  QB_CUDA(cudaMalloc((void**)&nl->enl_dev, sizeof(double)));
  QB_CUDA(cudaFuncSetAttribute(k_fnl<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, FNL_SMEM_BYTES));
  QB_CUDA(cudaFuncSetAttribute(k_fnl<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, FNL_SMEM_BYTES));
  QB_CUDA(cudaFuncSetAttribute(k_back<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, BK_SMEM_BYTES));
  QB_CUDA(cudaFuncSetAttribute(k_back<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, BK_SMEM_BYTES));
  *out = nl;
  return QB200_OK;
}

extern "C" int qb200_nl_add_species(qb200_nl* nl, int na, int npr, const int* lproj, const double* wt, const double* twnl,
                                    const double* tau)
{
  if (!nl || na < 0 || npr < 0) { set_error("qb200_nl_add_species: bad argument"); return QB200_EINVAL; }
  QB_CUDA(cudaSetDevice(nl->device));
  NlSpecies s;
  s.na = na; s.npr = npr; s.M = na * npr;
  s.lproj = nullptr; s.wt = nullptr; s.twnl = nullptr; s.tau = nullptr; s.ph = nullptr;
  if (s.M > 0) {
    if (!lproj || !wt || !twnl || !tau) { set_error("qb200_nl_add_species: null table"); return QB200_EINVAL; }
    for (int i = 0; i < npr; i++) if (lproj[i] < 0 || lproj[i] > 3) { set_error("qb200_nl_add_species: l > 3 unsupported (as in the reference)"); return QB200_EUNSUPPORTED; }
    int rc;
    if ((rc = nl_upload(nl, lproj, npr, &s.lproj)) || (rc = nl_upload(nl, wt, npr, &s.wt)) ||
        (rc = nl_upload(nl, twnl, (size_t)npr * nl->ngw, &s.twnl)) || (rc = nl_upload(nl, tau, 3 * (size_t)na, &s.tau))) return rc;
  }
  nl->sp.push_back(s);
  nl->ph.push_back(nullptr);
  nl->ph_dirty = true;
  return QB200_OK;
}

extern "C" int qb200_nl_set_lattice(qb200_nl* nl, const int* idx, const double* b, const double* kpoint)
{
  if (!nl || !idx || !b || !kpoint) { set_error("qb200_nl_set_lattice: bad argument"); return QB200_EINVAL; }
  QB_CUDA(cudaSetDevice(nl->device));
  const int ngw = nl->ngw;
  std::vector<int> planes(3 * (size_t)ngw);
  int jmax[3] = { 0, 0, 0 };
  for (int i = 0; i < ngw; i++)
    for (int d = 0; d < 3; d++) {
      const int v = idx[3 * (size_t)i + d];
      planes[(size_t)d * ngw + i] = v;
      jmax[d] = std::max(jmax[d], std::abs(v));
    }
  const int* dev;
  int rc = nl_upload(nl, planes.data(), planes.size(), &dev);
  if (rc) return rc;
  nl->lat.idx = dev;
  for (int d = 0; d < 3; d++) nl->lat.jmax[d] = jmax[d];
  nl->lat.JT = 2 * (jmax[0] + jmax[1] + jmax[2]) + 3;
  for (int i = 0; i < 9; i++) nl->bvec[i] = b[i];
  for (int d = 0; d < 3; d++) nl->kcart[d] = kpoint[0] * b[d] + kpoint[1] * b[3 + d] + kpoint[2] * b[6 + d];
  nl->ph_dirty = true;
  return QB200_OK;
}

// (re)build the separable phase tables after a change of positions or lattice
static int nl_refresh_tables(qb200_nl* nl)
{
  if (!nl->ph_dirty) return QB200_OK;
  nl->ph_dirty = false;
  if (!nl->lat.idx) return QB200_OK;
  for (size_t is = 0; is < nl->sp.size(); is++) {
    NlSpecies& S = nl->sp[is];
    if (S.M <= 0) continue;
    if (!nl->ph[is]) {
      QB_CUDA(cudaMalloc((void**)&nl->ph[is], (size_t)S.na * nl->lat.JT * sizeof(double2)));
      nl->owned.push_back(nl->ph[is]);
    }
    const double* b = nl->bvec;
    k_phase_tables<<<S.na, 128, 0, nl->stream>>>(S, nl->lat, b[0], b[1], b[2], b[3], b[4], b[5], b[6], b[7], b[8],
                                                 nl->kcart[0], nl->kcart[1], nl->kcart[2], nl->ph[is]);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return qb200::cuda_fail(e, "k_phase_tables launch", __FILE__, __LINE__);
    nl->launches++;
    S.ph = nl->ph[is];
  }
  return QB200_OK;
}

extern "C" int qb200_nl_set_positions(qb200_nl* nl, int is, const double* tau)
{
  if (!nl || is < 0 || is >= (int)nl->sp.size() || !tau) { set_error("qb200_nl_set_positions: bad argument"); return QB200_EINVAL; }
  QB_CUDA(cudaSetDevice(nl->device));
  if (nl->sp[is].na > 0 && nl->sp[is].tau)
    QB_CUDA(cudaMemcpy(const_cast<double*>(nl->sp[is].tau), tau, 3 * (size_t)nl->sp[is].na * sizeof(double), cudaMemcpyHostToDevice));
  nl->ph_dirty = true;
  return QB200_OK;
}

extern "C" int qb200_nl_set_stream(qb200_nl* nl, void* s)
{
  if (!nl) return QB200_EINVAL;
  nl->stream = (cudaStream_t)s;
  return QB200_OK;
}

extern "C" int qb200_nl_destroy(qb200_nl* nl)
{
  if (!nl) return QB200_OK;
  cudaSetDevice(nl->device);
  for (void* p : nl->owned) cudaFree(p);
  for (double* p : { nl->part, nl->fs, nl->eblk, nl->occ_dev, nl->enl_dev, nl->st_c, nl->st_cp }) if (p) cudaFree(p);
  delete nl;
  return QB200_OK;
}

extern "C" long long qb200_nl_query(const qb200_nl* nl, int what)
{
  if (!nl) return -1;
  if (what == 9) return nl->launches;
  return -1;
}

#define NL_LAUNCH_CHECK(nl) do { (nl)->launches++; cudaError_t e__ = cudaGetLastError(); \
    if (e__ != cudaSuccess) return qb200::cuda_fail(e__, "kernel launch", __FILE__, __LINE__); } while (0)

// device pointers; enl accumulated into nl->enl_dev (zeroed here)
int qb200_nl_energy_dev(qb200_nl* nl, int ldc, int nst, const double* c, const double* occ_host, int compute_hpsi, double* cp)
{
  int rc;
  if ((rc = nl_ensure(&nl->occ_dev, &nl->occ_cap, nst))) return rc;
  QB_CUDA(cudaMemcpyAsync(nl->occ_dev, occ_host, nst * sizeof(double), cudaMemcpyDefault, nl->stream));
  QB_CUDA(cudaMemsetAsync(nl->enl_dev, 0, sizeof(double), nl->stream));
  const int ncols = nl->is_real ? nst : 2 * nst;
  if ((rc = nl_refresh_tables(nl))) return rc;
  for (const NlSpecies& S : nl->sp) {
    if (S.M <= 0) continue;
    const int Mp = (S.M + 1) & ~1;               // even pitch: 16-byte cp.async chunks of fs stay aligned; the pad is zero
    const int PT = nl->is_real ? NL_TM : NL_TM / 2;
    const int mt = (S.M + PT - 1) / PT, nt = (nst + NL_TN - 1) / NL_TN;
    // split K so that the CTAs fill whole waves of the SMs (one CTA per SM); chunks are multiples of 16 plane waves
    int ksplit = 1;
    {
      const int maxk = std::max(1, nl->ngw / 512);
      double best = -1.0;
      for (int k = 1; k <= std::min(maxk, 64); k++) {
        const long ctas = (long)mt * nt * k;
        const long waves = (ctas + nl->nsm - 1) / nl->nsm;
        const double eff = (double)ctas / (double)(waves * nl->nsm) - 0.002 * k;   // mild preference for fewer partials
        if (eff > best) { best = eff; ksplit = k; }
      }
    }
    int gchunk = (nl->ngw + ksplit - 1) / ksplit;
    gchunk = ((gchunk + 15) / 16) * 16;
    ksplit = (nl->ngw + gchunk - 1) / gchunk;
    if ((rc = nl_ensure(&nl->part, &nl->part_cap, (size_t)ksplit * ncols * Mp))) return rc;
    if ((rc = nl_ensure(&nl->fs, &nl->fs_cap, 2 * (size_t)nst * Mp))) return rc;
    if (Mp != S.M) QB_CUDA(cudaMemsetAsync(nl->fs, 0, 2 * (size_t)nst * Mp * sizeof(double), nl->stream));
    const size_t total = (size_t)nst * S.M;
    const int nblk = (int)((total + 255) / 256);
    if ((rc = nl_ensure(&nl->eblk, &nl->eblk_cap, nblk))) return rc;
    dim3 g1(mt, nt, ksplit);
    prof_begin(3, nl->stream);
    if (nl->is_real) k_fnl<1><<<g1, NL_THREADS, FNL_SMEM_BYTES, nl->stream>>>(S, nl->lat, nl->ngw, nl->kpgx, (const double2*)c, ldc, nst, gchunk, nl->part, Mp);
    else k_fnl<0><<<g1, NL_THREADS, FNL_SMEM_BYTES, nl->stream>>>(S, nl->lat, nl->ngw, nl->kpgx, (const double2*)c, ldc, nst, gchunk, nl->part, Mp);
    prof_end(nl->stream);
    NL_LAUNCH_CHECK(nl);
    prof_begin(4, nl->stream);
    if (nl->is_real) k_fnl_finish<1><<<nblk, 256, 0, nl->stream>>>(S, nl->part, Mp, nst, ksplit, nl->occ_dev, 1.0 / nl->omega, nl->fs, nl->eblk);
    else k_fnl_finish<0><<<nblk, 256, 0, nl->stream>>>(S, nl->part, Mp, nst, ksplit, nl->occ_dev, 1.0 / nl->omega, nl->fs, nl->eblk);
    NL_LAUNCH_CHECK(nl);
    k_sum_blocks<<<1, 32, 0, nl->stream>>>(nl->eblk, nblk, nl->enl_dev);
    prof_end(nl->stream);
    NL_LAUNCH_CHECK(nl);
    if (compute_hpsi) {
      dim3 g2((nl->ngw + 63) / 64, (nst + NL_TN - 1) / NL_TN);
      prof_begin(5, nl->stream);
      if (nl->is_real) k_back<1><<<g2, NL_THREADS, BK_SMEM_BYTES, nl->stream>>>(S, nl->lat, nl->ngw, nl->kpgx, nl->fs, Mp, (double2*)cp, ldc, nst);
      else k_back<0><<<g2, NL_THREADS, BK_SMEM_BYTES, nl->stream>>>(S, nl->lat, nl->ngw, nl->kpgx, nl->fs, Mp, (double2*)cp, ldc, nst);
      prof_end(nl->stream);
      NL_LAUNCH_CHECK(nl);
    }
  }
  return QB200_OK;
}

extern "C" int qb200_nl_energy(qb200_nl* nl, int ldc, int nst, const double* c, const double* occ, int compute_hpsi,
                               double* cp, double* enl)
{
  if (!nl || !c || !occ || nst < 0 || ldc < nl->ngw || (compute_hpsi && !cp)) { set_error("qb200_nl_energy: bad argument"); return QB200_EINVAL; }
  if (enl) *enl = 0.0;
  if (nst == 0) return QB200_OK;
  QB_CUDA(cudaSetDevice(nl->device));
  const size_t blk = 2 * (size_t)ldc * nst;
  const double* cd = c;
  double* cpd = cp;
  int rc;
  if (!is_device_ptr(c)) {
    if ((rc = nl_ensure(&nl->st_c, &nl->st_c_cap, blk))) return rc;
    QB_CUDA(cudaMemcpyAsync(nl->st_c, c, blk * sizeof(double), cudaMemcpyHostToDevice, nl->stream));
    cd = nl->st_c;
  }
  if (compute_hpsi && !is_device_ptr(cp)) {
    if ((rc = nl_ensure(&nl->st_cp, &nl->st_cp_cap, blk))) return rc;
    QB_CUDA(cudaMemcpyAsync(nl->st_cp, cp, blk * sizeof(double), cudaMemcpyHostToDevice, nl->stream));
    cpd = nl->st_cp;
  }
  if ((rc = qb200_nl_energy_dev(nl, ldc, nst, cd, occ, compute_hpsi, cpd))) return rc;
  if (compute_hpsi && cpd != cp) QB_CUDA(cudaMemcpyAsync(cp, cpd, blk * sizeof(double), cudaMemcpyDeviceToHost, nl->stream));
  double e = 0.0;
  QB_CUDA(cudaMemcpyAsync(&e, nl->enl_dev, sizeof(double), cudaMemcpyDeviceToHost, nl->stream));
  QB_CUDA(cudaStreamSynchronize(nl->stream));
  if (enl) *enl = e;
  return QB200_OK;
}

double* qb200_nl_enl_dev(qb200_nl* nl) { return nl->enl_dev; }
cudaStream_t qb200_nl_swap_stream(qb200_nl* nl, cudaStream_t s) { cudaStream_t o = nl->stream; nl->stream = s; return o; }
