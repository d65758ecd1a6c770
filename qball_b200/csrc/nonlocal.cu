// qball_b200/csrc/nonlocal.cu -- NonLocalPotential::energy, norm-conserving branch
// (/root/reference/src/qball/NonLocalPotential.cc:1909-2171, 2628-2643) as two FP64 tensor-core GEMMs.
//
// The reference materialises anl[ig,(ia,ipr)] = twnl[ipr][ig] * (-i)^l * exp(-i (k+G).tau_ia)  (:1959-2036) for blocks
// of <=128 atoms and calls BLAS:  fnl = anl^H c (:2050-2068),  E_nl += occ wt/omega |fnl|^2 (:2106-2148),
// cp += anl (wt/omega fnl) (:2150-2171).  Here:
//   * k_anl_gen writes anl for ALL projectors (species concatenated, atom-major p = ia*npr + ipr within a species) and a
//     CHUNK of plane waves into one real matrix W in the form both GEMMs consume,
//        complex:  W[2p  ][2g..2g+1] = ( a.x, a.y)     W[2p+1][2g..2g+1] = (-a.y, a.x)        (a = anl[g,p])
//        Gamma:    W[p][2g..2g+1]    = ( a.x, a.y)     (a.x halved at G=0: the dger fix :2070-2082)
//     The chunk is the whole sphere when W fits the workspace (MgO216: 1.3 GB; written once per call like the
//     reference's comp_anl, or kept until the atoms move with QB200_ANL_CACHE=1); otherwise (Au992: 181 GB for
//     everything) the two sweeps below regenerate it chunk by chunk.
//   * k_fnl : part[(p,re|im), n] (+)= sum_k W[row][k] c[k, n]      k over the reals (g,re),(g,im) of the chunk
//   * k_back: cp[(g,re|im), n]  += sum_row W[row][(g,.)] fs[row, n]    (W used k-major: the same matrix, transposed role)
//     are pure cp.async -> shared memory -> mma.sync.m8n8k4.f64 (DMMA) pipelines, three stages deep, one barrier per
//     stage.  tcgen05 has no FP64 kind, so this is the tensor path there is for this contraction.
// Why anl is not generated inside the GEMMs (as an earlier version did): on B200 DMMA and plain FP64 instructions
// share one pipe (tools/microbench/dmma_peak.cu: 37.0 TF/s DMMA alone, 34 TF/s DFMA alone, the sum stays ~35 when
// mixed), and a warp that evaluates phase factors while others multiply finds its handful of DMUL/DFMA starved behind
// their DMMAs (measured: 4400 cycles for two DMULs) -- the tile generation could not be overlapped and cost 20 %.
// Summation is deterministic: split-K partials and chunks are reduced in fixed order, E_nl by a fixed tree.
#include "qb200_internal.h"
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>

namespace qb200 {

// GEMM tile geometry: CTA = 512 threads = 16 warps (4 x 4), warp tile 32 x 32 (16 m8n8 accumulators),
// CTA tile 128 x 128 reals, 32 reals of the reduction dimension per stage, 3 stages of shared memory.
#define NL_TM 128
#define NL_TN 128
#define NL_KSTEP 32
#define NL_THREADS 512
#define NL_NSTAGE 3
#define NL_PITCH 36            // doubles per row of a [row][k] tile: 32 + 4 -> conflict-free DMMA fragment loads
#define NL_PITCH_KR 132        // doubles per k-row of a [k][row] tile: 128 + 4 (same property, k-major)
#define NL_STAGE_RK (NL_TM * NL_PITCH + NL_TN * NL_PITCH)          // doubles per stage, A row-major (k_fnl)
#define NL_STAGE_KR (NL_KSTEP * NL_PITCH_KR + NL_TN * NL_PITCH)    // doubles per stage, A k-major (k_back)
#define FNL_SMEM_BYTES (NL_NSTAGE * NL_STAGE_RK * 8)
#define BK_SMEM_BYTES (NL_NSTAGE * NL_STAGE_KR * 8)

__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b)
{
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
// 16-byte asynchronous copy global -> shared; !valid copies nothing and writes zeros
__device__ __forceinline__ void nl_cp16(void* smem, const void* gmem, bool valid)
{
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  const int n = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(sa), "l"(gmem), "r"(n) : "memory");
}
__device__ __forceinline__ void nl_cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void nl_cp_wait_group() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

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
  int poff;                       // first projector of the species in the concatenated list
  const int* lproj;               // [npr]
  const double* wt;               // [npr]
  const double* twnl;             // [npr][ngw]
  const double* tau;              // [na][3]
  const double2* ph;              // [na][JT] separable phase tables (NlLattice), or null
};

// Optional integer description of the plane waves (qb200_nl_set_lattice): k+G = kpoint + h b0 + k b1 + l b2, so
//   exp(-i (k+G).tau) = [exp(-i kpoint.tau) exp(-i h b0.tau)] * exp(-i k b1.tau) * exp(-i l b2.tau)
// and the FP64 sincos per (atom, G) becomes three table look-ups and two complex multiplications.
// Tables: per atom JT = J0+J1+J2 entries, Jd = 2*jmax[d]+1.
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

// ------------------------------------------------------------------------------------------------ anl chunk -> W
// grid (ceil(gpad/128), na), block 128: one atom x 128 plane waves of the chunk [gbeg, gbeg+gcount), columns up to gpad
// (a multiple of 16) zero-filled.  W row pitch WP doubles.
template <int IS_REAL>
__global__ void __launch_bounds__(128) k_anl_gen(NlSpecies S, NlLattice L, int ngw, const double* __restrict__ kpgx, int gbeg,
                                                 int gcount, int gpad, double* __restrict__ W, size_t WP)
{
  const int gl = blockIdx.x * 128 + threadIdx.x;
  if (gl >= gpad) return;
  const int ia = blockIdx.y, g = gbeg + gl;
  const bool ok = gl < gcount;
  double sn = 0.0, cs = 0.0;
  if (ok) {
    if (L.idx != nullptr && S.ph != nullptr) {
      const double2 e = nl_phase(L, S.ph + (size_t)ia * L.JT, L.idx[g], L.idx[(size_t)ngw + g], L.idx[2 * (size_t)ngw + g]);
      cs = e.x; sn = e.y;
    } else {
      const double arg = -(kpgx[g] * S.tau[3 * ia] + kpgx[(size_t)ngw + g] * S.tau[3 * ia + 1] + kpgx[2 * (size_t)ngw + g] * S.tau[3 * ia + 2]);
      sincos(arg, &sn, &cs);
    }
  }
  for (int ipr = 0; ipr < S.npr; ipr++) {
    const size_t p = (size_t)S.poff + (size_t)ia * S.npr + ipr;
    double2 a = make_double2(0.0, 0.0);
    if (ok) {
      a = anl_value(S.lproj[ipr], S.twnl[(size_t)ipr * ngw + g], sn, cs);
      if (IS_REAL && g == 0) a.x *= 0.5;           // G=0 counted once: dger fix, NonLocalPotential.cc:2078-2080
    }
    if (IS_REAL) {
      *reinterpret_cast<double2*>(W + p * WP + 2 * gl) = a;
    } else {
      *reinterpret_cast<double2*>(W + (2 * p) * WP + 2 * gl) = a;
      *reinterpret_cast<double2*>(W + (2 * p + 1) * WP + 2 * gl) = make_double2(-a.y, a.x);
    }
  }
}

// one stage of the warp tile: acc[i][j] += A(32 x 32 reals) * B(32 x 32 reals)
// k4 steps [K4A, K4B) of one stage (the whole stage by default): the kernels run the first steps, then issue the copies of the
// stage after next, then the rest -- right after the stage barrier the tensor pipe gets work at once instead of idling while
// all 16 warps compute copy addresses (ncu: 11 % of the samples of k_fnl<1> sat behind that barrier)
template <bool A_KMAJOR, int K4A = 0, int K4B = NL_KSTEP / 4>
__device__ __forceinline__ void warp_mma_stage(const double* __restrict__ As, const double* __restrict__ Bs, double (&acc)[4][4][2],
                                               int lane, int wm, int wn)
{
  const int r = lane >> 2, kq = lane & 3;
  const double* a0 = A_KMAJOR ? As + kq * NL_PITCH_KR + wm * 32 + r : As + (wm * 32 + r) * NL_PITCH + kq;
  const double* b0 = Bs + (wn * 32 + r) * NL_PITCH + kq;
#pragma unroll
  for (int k4 = K4A; k4 < K4B; k4++) {
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

// ------------------------------------------------------------------------------------------------ fnl = anl^H c
// grid (ceil(RW/128), ceil(nst/128), ksplit).  Rows = rows of W (complex: (p,re),(p,im); Gamma: p), columns = states,
// reduction over the chunk's reals k = 2*(g-gbeg)+{re,im}; this CTA takes [blockIdx.z*kper, +kper) of them.
// part[(ks*ncols + col)*Mp + p] (=, or += when accumulate), ncols = IS_REAL ? nst : 2*nst, col = n or 2n+{re,im}.
template <int IS_REAL>
__global__ void __launch_bounds__(NL_THREADS, 1) k_fnl(const double* __restrict__ W, size_t WP, int RW, int gbeg, int gcount,
                                                         int kper, const double2* __restrict__ c, size_t ldc, int nst,
                                                         double* __restrict__ part, int Mp, int Mtot, int accumulate,
                                                         int tn = NL_TN)
{
  // tn (a multiple of 32, <= NL_TN): columns per CTA.  A block whose column count is not a multiple of 128 is cut into EQUAL
  // tiles (192 columns: 2 x 96 instead of 128 + 64), so every CTA keeps the same number of warps busy
  extern __shared__ __align__(16) double nl_smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // warp tile (wm, wn) of the 4 x 4: skewed so that the warps of one tile row AND those of one tile column sit on four different
  // SM sub-partitions (warp % 4) -- when a row or a column of warp tiles has no work (padding), every FP64 tensor pipe keeps
  // three busy warps instead of one pipe idling
  const int wm = warp >> 2, wn = (warp + wm) & 3;
  const int r0 = blockIdx.x * NL_TM, n0 = blockIdx.y * tn;
  const int nend = min(nst, n0 + tn);
  const int kbeg = blockIdx.z * kper, kend = min(kbeg + kper, 2 * gcount);
  const int nstage = (kend - kbeg + NL_KSTEP - 1) / NL_KSTEP;
  // this thread's 4+4 copies per stage: row (tid>>4) + 32 i, 16-byte chunk (tid & 15) of the stage's 32 reals
  const int crow = tid >> 4, cch = tid & 15;
  auto issue = [&](int st) {
    if (st < nstage) {
      double* As = nl_smem + (st % NL_NSTAGE) * NL_STAGE_RK;
      double* Bs = As + NL_TM * NL_PITCH;
      const int k = kbeg + st * NL_KSTEP + 2 * cch;               // chunk-local real index of this copy
      const bool kok = k < kend;
#pragma unroll
      for (int i = 0; i < 4; i++) {
        const int row = crow + 32 * i;
        const bool aok = kok && r0 + row < RW;
        nl_cp16(As + row * NL_PITCH + 2 * cch, W + (aok ? (size_t)(r0 + row) * WP + k : 0), aok);
        const bool bok = kok && n0 + row < nend;
        nl_cp16(Bs + row * NL_PITCH + 2 * cch, c + (bok ? (size_t)(n0 + row) * ldc + gbeg + (k >> 1) : 0), bok);
      }
    }
    nl_cp_commit();
  };
  double acc[4][4][2];
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) acc[i][j][0] = acc[i][j][1] = 0.0;
  for (int s = 0; s < NL_NSTAGE - 1; s++) issue(s);
  for (int st = 0; st < nstage; st++) {
    nl_cp_wait_group<NL_NSTAGE - 2>();
    __syncthreads();                   // stage st has landed for everybody; everybody is done with stage st-1's buffer
    const double* As = nl_smem + (st % NL_NSTAGE) * NL_STAGE_RK;
    // the last row tile is mostly padding when RW is not a multiple of 128 (MgO216: 540 rows, 28 of 128 in the fifth
    // tile): warps whose 32-row slab holds no row skip the tensor work, the others then own the pipe (warp-uniform test)
    const bool work = r0 + wm * 32 < RW && n0 + wn * 32 < nend;
    if (work) warp_mma_stage<false, 0, 2>(As, As + NL_TM * NL_PITCH, acc, lane, wm, wn);
    issue(st + NL_NSTAGE - 1);
    if (work) warp_mma_stage<false, 2, NL_KSTEP / 4>(As, As + NL_TM * NL_PITCH, acc, lane, wm, wn);
  }
  const int ncols = IS_REAL ? nst : 2 * nst;
  const int r = lane >> 2, cq = lane & 3;
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++)
#pragma unroll
      for (int e = 0; e < 2; e++) {
        const int row = r0 + wm * 32 + i * 8 + r;
        const int n = n0 + wn * 32 + j * 8 + 2 * cq + e;
        const int p = IS_REAL ? row : (row >> 1);
        const int col = IS_REAL ? n : 2 * n + (row & 1);
        if (p < Mtot && n < nend) {
          double* dst = part + ((size_t)blockIdx.z * ncols + col) * Mp + p;
          *dst = accumulate ? *dst + acc[i][j][e] : acc[i][j][e];
        }
      }
}

// ------------------------------------------------------------------------------------------------ E_nl, fnl <- wt/omega fnl
// one thread per (n, p); fs[n][p] complex (or real at Gamma); block partial sums of E_nl to eblk
template <int IS_REAL>
__global__ void __launch_bounds__(256) k_fnl_finish(const double* __restrict__ wtp, int Mtot, const double* __restrict__ part, int Mp,
                                                    int nst, int ksplit, const double* __restrict__ occ, double omega_inv,
                                                    double* __restrict__ fs, double* __restrict__ eblk)
{
  __shared__ double red[256];
  const size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  const size_t total = (size_t)nst * Mtot;
  double e = 0.0;
  if (idx < total) {
    const int n = (int)(idx / Mtot), p = (int)(idx % Mtot);
    const int ncols = IS_REAL ? nst : 2 * nst;
    const double fac = wtp[p] * omega_inv;
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

// *acc += sum of eblk[0..n): one block of 256 threads, fixed assignment and fixed tree -> deterministic
__global__ void __launch_bounds__(256) k_sum_blocks(const double* __restrict__ eblk, int n, double* __restrict__ acc)
{
  __shared__ double red[256];
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += 256) s += eblk[i];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int k = 128; k > 0; k >>= 1) {
    if (threadIdx.x < k) red[threadIdx.x] += red[threadIdx.x + k];
    __syncthreads();
  }
  if (threadIdx.x == 0) *acc += red[0];
}

// ------------------------------------------------------------------------------------------------ cp += anl * fs
// grid (ceil(nst/128), ceil(gcount/64)) -- state tiles fastest, so the CTAs that share a W tile run together and W streams
// from HBM once (r1l ncu: 4.7 GB per launch with the plane-wave tiles fastest, W re-read by every state tile):
// 128 output reals (64 plane waves of the chunk x re/im) x 128 states per CTA.
// Reduction over the rows of W (32 per stage), read k-major: A[k = W row][row = 2*(g-gbeg)+{re,im}];
// B[k][n] = fs[n][k], fs = wt/omega * fnl with row pitch FP = (IS_REAL ? Mp : 2*Mp) doubles, zero beyond RW.
// IS_REAL == 2: the half-sphere form of complex states at Gamma (see k_split_pm): columns (2n, 2n+1) carry (f_r, f_i) of state
// n, rows (2g', 2g'+1) the (A, B) parts; the epilogue forms cp(G) (+)= (P - Q) + i (R + T), cp(-G) (+)= (P + Q) + i (R - T)
// with P = sum A f_r, T = sum B f_r, R = sum A f_i, Q = sum B f_i directly from the accumulators (one lane exchange
// between the A-row and the B-row of a plane wave) and stores them at ghalf[g'] / gminus[g'] of the sphere-ordered block.
template <int IS_REAL>
__global__ void __launch_bounds__(NL_THREADS, 1) k_back(const double* __restrict__ W, size_t WP, int RW, int gbeg, int gcount,
                                                          const double* __restrict__ fs, int FP, double2* __restrict__ cp,
                                                          size_t ldc, int nst, int overwrite, const int* __restrict__ ghalf = nullptr,
                                                          const int* __restrict__ gminus = nullptr, int tn = NL_TN)
{
  extern __shared__ __align__(16) double nl_smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp >> 2, wn = (warp + wm) & 3;                // skewed as in k_fnl
  const int gl0 = blockIdx.y * 64, n0 = blockIdx.x * tn;         // tn: equal column tiles, as in k_fnl
  const int nend = min(nst, n0 + tn);
  const bool work = n0 + wn * 32 < nend;                         // (warp-uniform) this warp's 32 columns hold a state
  const int nstage = (RW + NL_KSTEP - 1) / NL_KSTEP;
  auto issue = [&](int st) {
    if (st < nstage) {
      double* As = nl_smem + (st % NL_NSTAGE) * NL_STAGE_KR;
      double* Bs = As + NL_KSTEP * NL_PITCH_KR;
      const int k0 = st * NL_KSTEP;
#pragma unroll
      for (int i = 0; i < 4; i++) {
        // A: 32 k-rows x 64 chunks (one plane wave = (re,im) each)
        const int ci = tid + i * NL_THREADS, kr = ci >> 6, gc = ci & 63;
        const bool aok = k0 + kr < RW && gl0 + gc < gcount;
        nl_cp16(As + kr * NL_PITCH_KR + 2 * gc, W + (aok ? (size_t)(k0 + kr) * WP + 2 * (size_t)(gl0 + gc) : 0), aok);
        // B: 128 states x 16 chunks of 2 reals
        const int nl = ci >> 4, ch = ci & 15;
        const bool bok = n0 + nl < nend && k0 + 2 * ch < FP;
        nl_cp16(Bs + nl * NL_PITCH + 2 * ch, fs + (bok ? (size_t)(n0 + nl) * FP + k0 + 2 * ch : 0), bok);
      }
    }
    nl_cp_commit();
  };
  double acc[4][4][2];
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) acc[i][j][0] = acc[i][j][1] = 0.0;
  for (int s = 0; s < NL_NSTAGE - 1; s++) issue(s);
  for (int st = 0; st < nstage; st++) {
    nl_cp_wait_group<NL_NSTAGE - 2>();
    __syncthreads();
    const double* As = nl_smem + (st % NL_NSTAGE) * NL_STAGE_KR;
    if (work) warp_mma_stage<true, 0, 2>(As, As + NL_KSTEP * NL_PITCH_KR, acc, lane, wm, wn);
    issue(st + NL_NSTAGE - 1);
    if (work) warp_mma_stage<true, 2, NL_KSTEP / 4>(As, As + NL_KSTEP * NL_PITCH_KR, acc, lane, wm, wn);
  }
  const int r = lane >> 2, cq = lane & 3;
  if (IS_REAL == 2) {
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const int row = wm * 32 + i * 8 + r;                   // even: A-row (P, R), odd: B-row (T, Q) of plane wave gl
      const int gl = gl0 + (row >> 1);
      const bool odd = row & 1;
      const int gdst = gl < gcount ? (odd ? gminus[gl] : ghalf[gl]) : 0;
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const double mine0 = acc[i][j][0], mine1 = acc[i][j][1];
        const double oth0 = __shfl_xor_sync(0xffffffffu, mine0, 4), oth1 = __shfl_xor_sync(0xffffffffu, mine1, 4);
        const double Pv = odd ? oth0 : mine0, Rv = odd ? oth1 : mine1, Tv = odd ? mine0 : oth0, Qv = odd ? mine1 : oth1;
        const int n2 = n0 + wn * 32 + j * 8 + 2 * cq;          // column pair (2n, 2n+1) of state n = n2 / 2
        if (gl < gcount && n2 < nend && !(odd && gl == 0)) {  // G = 0 (gl == 0) has no partner
          double2* dst = cp + (size_t)(n2 >> 1) * ldc + gdst;
          double2 v = overwrite ? make_double2(0.0, 0.0) : *dst;
          if (odd) { v.x += Pv + Qv; v.y += Rv - Tv; } else { v.x += Pv - Qv; v.y += Rv + Tv; }
          *dst = v;
        }
      }
    }
    return;
  }
  double* cpd = reinterpret_cast<double*>(cp);
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++)
#pragma unroll
      for (int e = 0; e < 2; e++) {
        const int row = wm * 32 + i * 8 + r;                 // real row within the tile: 2*gl + (re|im)
        const int gl = gl0 + (row >> 1);
        const int g = gbeg + gl;
        const int n = n0 + wn * 32 + j * 8 + 2 * cq + e;
        double v = acc[i][j][e];
        if (IS_REAL == 1 && g == 0 && (row & 1) == 0) v *= 2.0;    // W holds half of Re anl at G=0 (k_anl_gen)
        if (gl < gcount && n < nend) {
          double* dst = cpd + 2 * ((size_t)n * ldc + g) + (row & 1);
          *dst = overwrite ? v : *dst + v;
        }
      }
}

// ------------------------------------------------------------------------------------------------ Gamma point, complex states
// Complex wavefunctions at k = 0 (force_complex_wf, the TDDFT configuration; MgO216): the projectors are transforms of
// REAL functions, anl(-G) = conj(anl(G)), and the sphere holds every G together with -G.  With a = c(G), b = c(-G),
// anl(G) = A + iB:   conj(anl(G)) a + conj(anl(-G)) b = A (a+b) - iB (a-b), so over the half sphere G > 0 (G = 0 as an
// unpaired entry, b = 0)
//   Re fnl = sum_G [A,B].[Re(a+b),  Im(a-b)]          Im fnl = sum_G [A,B].[Im(a+b), -Re(a-b)]
// -- two REAL dot products with the same real row (A_G, B_G): psi = psi_R + i psi_I is projected as two real functions,
// 2 real MACs per (projector, state, plane wave) instead of the 3 of the Karatsuba form.  The back-projection is the
// transposed statement.  k_split_pm builds the half-sphere block, the epilogue of k_back<2> scatters back to the sphere; the GEMMs are the
// real-basis kernels k_fnl<1> / k_back on 2*nst real "states" and ngw reals (= 2 * half sphere) per state.
// grid (ceil(hpad/128), na): anl of the half sphere, W[p][2g'..2g'+1] = (A, B), g = ghalf[g']; columns >= nhalf zero
__global__ void __launch_bounds__(128) k_anl_gen_half(NlSpecies S, NlLattice L, int ngw, const double* __restrict__ kpgx,
                                                      const int* __restrict__ ghalf, int nhalf, int hpad, double* __restrict__ W, size_t WP)
{
  const int gl = blockIdx.x * 128 + threadIdx.x;
  if (gl >= hpad) return;
  const int ia = blockIdx.y;
  const bool ok = gl < nhalf;
  const int g = ok ? ghalf[gl] : 0;
  double sn = 0.0, cs = 0.0;
  if (ok) {
    if (L.idx != nullptr && S.ph != nullptr) {
      const double2 e = nl_phase(L, S.ph + (size_t)ia * L.JT, L.idx[g], L.idx[(size_t)ngw + g], L.idx[2 * (size_t)ngw + g]);
      cs = e.x; sn = e.y;
    } else {
      const double arg = -(kpgx[g] * S.tau[3 * ia] + kpgx[(size_t)ngw + g] * S.tau[3 * ia + 1] + kpgx[2 * (size_t)ngw + g] * S.tau[3 * ia + 2]);
      sincos(arg, &sn, &cs);
    }
  }
  for (int ipr = 0; ipr < S.npr; ipr++) {
    const size_t p = (size_t)S.poff + (size_t)ia * S.npr + ipr;
    double2 a = make_double2(0.0, 0.0);
    if (ok) a = anl_value(S.lproj[ipr], S.twnl[(size_t)ipr * ngw + g], sn, cs);
    *reinterpret_cast<double2*>(W + p * WP + 2 * gl) = a;
  }
}
// The half-sphere form needs twnl(-G) = (-1)^l twnl(G) (true of the reference's tables, NonLocalPotential.cc:261-1522: real
// spherical harmonics times a radial function); the ABI accepts any table, so the property is CHECKED on the device
// whenever the tables change and the general path is used if it does not hold.
// grid (ceil(nhalf/256), npr): out[0] = max |twnl(-G) - (-1)^l twnl(G)|, out[1] = max |twnl| (non-negative doubles
// compare like their bit patterns)
__global__ void __launch_bounds__(256) k_twnl_parity(NlSpecies S, int ngw, const int* __restrict__ ghalf, const int* __restrict__ gminus,
                                                     int nhalf, unsigned long long* __restrict__ out)
{
  const int gl = blockIdx.x * 256 + threadIdx.x;
  double dev = 0.0, mag = 0.0;
  if (gl < nhalf) {
    const int ipr = blockIdx.y;
    const double tp = S.twnl[(size_t)ipr * ngw + ghalf[gl]], tm = S.twnl[(size_t)ipr * ngw + gminus[gl]];
    const double sgn = (S.lproj[ipr] & 1) ? -1.0 : 1.0;
    dev = fabs(tm - sgn * tp);
    mag = fmax(fabs(tp), fabs(tm));
    if (gl == 0 && (S.lproj[ipr] & 1)) dev = fabs(tp);        // odd l at G = 0 must vanish
  }
  for (int o = 16; o > 0; o >>= 1) {
    dev = fmax(dev, __shfl_xor_sync(0xffffffffu, dev, o));
    mag = fmax(mag, __shfl_xor_sync(0xffffffffu, mag, o));
  }
  if ((threadIdx.x & 31) == 0) {
    atomicMax(out, (unsigned long long)__double_as_longlong(dev));
    atomicMax(out + 1, (unsigned long long)__double_as_longlong(mag));
  }
}
// grid (ceil(nhalf/256), nst): U[2n][g'] = (Re(a+b), Im(a-b)), U[2n+1][g'] = (Im(a+b), -Re(a-b)); g' = 0 is G = 0 (b = 0)
__global__ void __launch_bounds__(256) k_split_pm(const double2* __restrict__ c, size_t ldc, const int* __restrict__ ghalf,
                                                  const int* __restrict__ gminus, int nhalf, double2* __restrict__ U, size_t ldu)
{
  const int gl = blockIdx.x * 256 + threadIdx.x;
  if (gl >= nhalf) return;
  const size_t n = blockIdx.y;
  const double2 a = c[n * ldc + ghalf[gl]];
  double2 b = make_double2(0.0, 0.0);
  if (gl > 0) b = c[n * ldc + gminus[gl]];
  U[(2 * n) * ldu + gl] = make_double2(a.x + b.x, a.y - b.y);
  U[(2 * n + 1) * ldu + gl] = make_double2(a.y + b.y, b.x - a.x);
}
// one thread per (n, p): split-K reduce of the 2*nst real columns, E_nl partials, fs[2n][p] = wt/omega Re fnl, fs[2n+1][p] = .. Im
__global__ void __launch_bounds__(256) k_fnl_finish_half(const double* __restrict__ wtp, int Mtot, const double* __restrict__ part, int Mp,
                                                         int nst, int ksplit, const double* __restrict__ occ, double omega_inv,
                                                         double* __restrict__ fs, double* __restrict__ eblk)
{
  __shared__ double red[256];
  const size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  const size_t total = (size_t)nst * Mtot;
  double e = 0.0;
  if (idx < total) {
    const int n = (int)(idx / Mtot), p = (int)(idx % Mtot);
    const int ncols = 2 * nst;
    const double fac = wtp[p] * omega_inv;
    double fr = 0.0, fi = 0.0;
    for (int ks = 0; ks < ksplit; ks++) {
      fr += part[((size_t)ks * ncols + 2 * n) * Mp + p];
      fi += part[((size_t)ks * ncols + 2 * n + 1) * Mp + p];
    }
    e = fac * occ[n] * (fr * fr + fi * fi);
    fs[(size_t)(2 * n) * Mp + p] = fac * fr;
    fs[(size_t)(2 * n + 1) * Mp + p] = fac * fi;
  }
  red[threadIdx.x] = e;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) eblk[blockIdx.x] = red[0];
}

}  // namespace qb200

#include "nonlocal_3m.cuh"

using namespace qb200;

// augmentation tables of the ultrasoft entry points (ultrasoft.cuh)
struct UsSpeciesDev {
  int nq;
  const int *lm1, *lm2;      // [nq] channels of pair q
  const double* dzero;       // [nq] D^0
  const double2* qnm;        // [nq][ngv] Q_q(G) on the density basis
};
struct UsTables {
  int ngv = 0;
  const double* vkpgx = nullptr;                 // device [3][ngv]
  std::vector<UsSpeciesDev> sp;                  // per species of the object (nq == 0: not set)
};

struct qb200_nl {
  UsTables* us;
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
  std::vector<int> ph_JT;                      // the NlLattice::JT each table was allocated for
  std::vector<void*> lat_owned;                // uploads of the last qb200_nl_set_lattice (freed by the next one)
  bool ph_dirty;
  // concatenated projector list and the materialised anl chunk
  int Mtot;
  double* wtp; size_t wtp_cap;                 // [Mtot] weight of every projector
  bool wtp_dirty;
  double* W; size_t W_cap;                     // anl chunk, RW rows x WP doubles
  long long anl_budget;                        // bytes W may take
  bool W_valid;                                // W holds the whole sphere for the current positions
  bool cache_anl;                              // keep a whole-sphere W between energy calls until the atoms move
  int nchunks_last;
  bool use3m;                                  // complex bases: Karatsuba 3-GEMM form (nonlocal_3m.cuh)
  int tile3m;                                  // 0: 512-thread CTAs (one per SM), 1: 256-thread CTAs (two per SM)
  size_t W_WP;                                 // row pitch W was last zero-filled for (3M pad rows must be zero)
  // Gamma point with complex states (k = 0, force_complex_wf): real-function split over the half sphere
  bool gamma_half;                             // available: lattice description given, k = 0, every G has its -G
  bool gamma_off;                              // QB200_NL_GAMMA=0
  int nhalf;                                   // (ngw + 1) / 2: G = 0 first, then one of every (G, -G) pair
  const int *ghalf, *gminus;                   // device [nhalf]: index of G and of -G in the basis order
  double *Wg, *Ug; size_t Wg_cap, Ug_cap;
  bool Wg_valid;
  bool sym_dirty, sym_ok;                      // twnl(-G) = (-1)^l twnl(G) verified for the current tables
  int last_mode;                               // projector path of the last energy call: 0 real basis, 1 four-product, 2 three-product, 3 Gamma half sphere
};

static int nl_ensure(double** buf, size_t* cap, size_t elems)
{
  if (*cap >= elems && *buf) return QB200_OK;
  if (*buf) { cudaFree(*buf); *buf = nullptr; *cap = 0; }
  QB_CUDA(cudaMalloc((void**)buf, std::max<size_t>(elems, 1) * sizeof(double)));
  *cap = elems;
  return QB200_OK;
}

template <class T> static int nl_upload(qb200_nl* nl, const T* h, size_t n, const T** d, bool lattice = false)
{
  void* p = nullptr;
  QB_CUDA(cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(T)));
  (lattice ? nl->lat_owned : nl->owned).push_back(p);
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
  nl->us = nullptr;
  nl->device = device; nl->stream = 0; nl->ngw = ngw; nl->is_real = is_real ? 1 : 0; nl->omega = omega;
  nl->part = nl->fs = nl->eblk = nl->occ_dev = nl->enl_dev = nl->st_c = nl->st_cp = nullptr;
  nl->part_cap = nl->fs_cap = nl->eblk_cap = nl->occ_cap = nl->st_c_cap = nl->st_cp_cap = 0;
  nl->launches = 0;
  nl->lat.idx = nullptr; nl->lat.jmax[0] = nl->lat.jmax[1] = nl->lat.jmax[2] = 0; nl->lat.JT = 0;
  nl->ph_dirty = true;
  nl->Mtot = 0; nl->wtp = nullptr; nl->wtp_cap = 0; nl->wtp_dirty = true;
  nl->W = nullptr; nl->W_cap = 0; nl->W_valid = false; nl->nchunks_last = 0;
  nl->anl_budget = 8ll << 30;
  if (const char* e = getenv("QB200_ANL_BYTES")) nl->anl_budget = std::max(1ll << 20, atoll(e));
  // a whole-sphere anl block stays valid until positions, tables, lattice or workspace change (each of those calls
  // invalidates it): the next energy call reuses it instead of regenerating it (QB200_ANL_CACHE=0: regenerate every call,
  // as the reference's comp_anl does)
  nl->cache_anl = true;
  if (const char* e = getenv("QB200_ANL_CACHE")) nl->cache_anl = e[0] != '0';
  nl->use3m = !is_real;
  if (const char* e = getenv("QB200_NL_3M")) if (e[0] == '0') nl->use3m = false;
  nl->tile3m = 1; nl->W_WP = 0;
  nl->gamma_half = false; nl->gamma_off = false; nl->nhalf = 0; nl->ghalf = nl->gminus = nullptr;
  nl->Wg = nl->Ug = nullptr; nl->Wg_cap = nl->Ug_cap = 0; nl->Wg_valid = false; nl->last_mode = 0; nl->sym_dirty = true; nl->sym_ok = false;
  if (const char* e = getenv("QB200_NL_GAMMA")) if (e[0] == '0') nl->gamma_off = true;
  if (const char* e = getenv("QB200_NL_TILE")) nl->tile3m = atoi(e);
  // every failure after `new` destroys the object again (no leak on an early return)
  auto fail = [nl](cudaError_t e, const char* what, int line) { const int rc = cuda_fail(e, what, __FILE__, line); qb200_nl_destroy(nl); return rc; };
#define NL_CREATE_CUDA(x) do { cudaError_t e__ = (x); if (e__ != cudaSuccess) return fail(e__, #x, __LINE__); } while (0)
  cudaDeviceProp prop;
  NL_CREATE_CUDA(cudaGetDeviceProperties(&prop, device));
  nl->nsm = prop.multiProcessorCount;
  const double* d;
  int rc = nl_upload(nl, kpgx, 3 * (size_t)ngw, &d);
  if (rc) { qb200_nl_destroy(nl); return rc; }
  nl->kpgx = const_cast<double*>(d);
  NL_CREATE_CUDA(cudaMalloc((void**)&nl->enl_dev, sizeof(double)));
  NL_CREATE_CUDA(cudaFuncSetAttribute(k_fnl<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, FNL_SMEM_BYTES));
  NL_CREATE_CUDA(cudaFuncSetAttribute(k_fnl<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, FNL_SMEM_BYTES));
  NL_CREATE_CUDA(cudaFuncSetAttribute(k_back<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, BK_SMEM_BYTES));
  NL_CREATE_CUDA(cudaFuncSetAttribute(k_back<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, BK_SMEM_BYTES));
  NL_CREATE_CUDA(cudaFuncSetAttribute(k_back<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, BK_SMEM_BYTES));
  NL_CREATE_CUDA(cudaFuncSetAttribute(k_fnl3<4, 4, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, Fnl3Cfg<4, 4, 3>::SMEM));
  NL_CREATE_CUDA(cudaFuncSetAttribute(k_fnl3<4, 2, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, Fnl3Cfg<4, 2, 2>::SMEM));
  NL_CREATE_CUDA(cudaFuncSetAttribute(k_back3<4, 4, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, Back3Cfg<4, 4, 4>::SMEM));
  NL_CREATE_CUDA(cudaFuncSetAttribute(k_back3<4, 2, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, Back3Cfg<4, 2, 3>::SMEM));
#undef NL_CREATE_CUDA
  *out = nl;
  return QB200_OK;
}

extern "C" int qb200_nl_add_species(qb200_nl* nl, int na, int npr, const int* lproj, const double* wt, const double* twnl,
                                    const double* tau)
{
  if (!nl || na < 0 || npr < 0) { set_error("qb200_nl_add_species: bad argument"); return QB200_EINVAL; }
  QB_CUDA(cudaSetDevice(nl->device));
  NlSpecies s;
  s.na = na; s.npr = npr; s.M = na * npr; s.poff = nl->Mtot;
  s.lproj = nullptr; s.wt = nullptr; s.twnl = nullptr; s.tau = nullptr; s.ph = nullptr;
  if (s.M > 0) {
    if (!lproj || !wt || !tau) { set_error("qb200_nl_add_species: null table"); return QB200_EINVAL; }
    for (int i = 0; i < npr; i++) if (lproj[i] < 0 || lproj[i] > 3) { set_error("qb200_nl_add_species: l > 3 unsupported (as in the reference)"); return QB200_EUNSUPPORTED; }
    int rc;
    if ((rc = nl_upload(nl, lproj, npr, &s.lproj)) || (rc = nl_upload(nl, wt, npr, &s.wt)) || (rc = nl_upload(nl, tau, 3 * (size_t)na, &s.tau))) return rc;
    if (twnl) { if ((rc = nl_upload(nl, twnl, (size_t)npr * nl->ngw, &s.twnl))) return rc; }
    else {                                     // no table yet: qb200_nl_update_twnl fills it on the device
      void* t = nullptr;
      QB_CUDA(cudaMalloc(&t, (size_t)npr * nl->ngw * sizeof(double)));
      nl->owned.push_back(t);
      QB_CUDA(cudaMemset(t, 0, (size_t)npr * nl->ngw * sizeof(double)));
      s.twnl = (const double*)t;
    }
  }
  nl->sp.push_back(s);
  nl->ph.push_back(nullptr);
  nl->ph_JT.push_back(0);
  nl->Mtot += s.M;
  nl->ph_dirty = true; nl->wtp_dirty = true; nl->W_valid = false; nl->Wg_valid = false; nl->sym_dirty = true;
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
  // a repeated call (cell or cutoff change) replaces the previous description: release its uploads first
  QB_CUDA(cudaStreamSynchronize(nl->stream));
  for (void* q : nl->lat_owned) cudaFree(q);
  nl->lat_owned.clear();
  nl->lat.idx = nullptr; nl->ghalf = nl->gminus = nullptr; nl->nhalf = 0;
  const int* dev;
  int rc = nl_upload(nl, planes.data(), planes.size(), &dev, true);
  if (rc) return rc;
  nl->lat.idx = dev;
  for (int d = 0; d < 3; d++) nl->lat.jmax[d] = jmax[d];
  nl->lat.JT = 2 * (jmax[0] + jmax[1] + jmax[2]) + 3;
  for (int i = 0; i < 9; i++) nl->bvec[i] = b[i];
  for (int d = 0; d < 3; d++) nl->kcart[d] = kpoint[0] * b[d] + kpoint[1] * b[3 + d] + kpoint[2] * b[6 + d];
  nl->ph_dirty = true; nl->W_valid = false; nl->Wg_valid = false; nl->sym_dirty = true;
  // Gamma point with complex states: pair every G with -G (possible iff k = 0 and the sphere is symmetric)
  nl->gamma_half = false;
  if (!nl->is_real && kpoint[0] == 0.0 && kpoint[1] == 0.0 && kpoint[2] == 0.0 && (ngw & 1)) {
    const long long J0 = 2ll * jmax[0] + 1, J1 = 2ll * jmax[1] + 1, J2 = 2ll * jmax[2] + 1;
    if (J0 * J1 * J2 < (1ll << 31)) {
      std::vector<int> where((size_t)(J0 * J1 * J2), -1);
      auto key = [&](int h, int k, int l) { return (size_t)(((long long)(h + jmax[0]) * J1 + (k + jmax[1])) * J2 + (l + jmax[2])); };
      for (int i = 0; i < ngw; i++) where[key(idx[3 * (size_t)i], idx[3 * (size_t)i + 1], idx[3 * (size_t)i + 2])] = i;
      std::vector<int> gh, gm;
      gh.reserve(ngw / 2 + 1); gm.reserve(ngw / 2 + 1);
      bool ok = where[key(0, 0, 0)] >= 0;
      if (ok) { gh.push_back(where[key(0, 0, 0)]); gm.push_back(where[key(0, 0, 0)]); }
      for (int i = 0; i < ngw && ok; i++) {
        const int h = idx[3 * (size_t)i], k = idx[3 * (size_t)i + 1], l = idx[3 * (size_t)i + 2];
        if (!(h > 0 || (h == 0 && (k > 0 || (k == 0 && l > 0))))) continue;       // one representative per pair
        const int m = where[key(-h, -k, -l)];
        if (m < 0) { ok = false; break; }
        gh.push_back(i); gm.push_back(m);
      }
      if (ok && (int)gh.size() == (ngw + 1) / 2) {
        const int *dh, *dm;
        if ((rc = nl_upload(nl, gh.data(), gh.size(), &dh, true)) || (rc = nl_upload(nl, gm.data(), gm.size(), &dm, true))) return rc;
        nl->ghalf = dh; nl->gminus = dm; nl->nhalf = (int)gh.size(); nl->gamma_half = true;
      }
    }
  }
  return QB200_OK;
}

#define NL_LAUNCH_CHECK(nl) do { (nl)->launches++; cudaError_t e__ = cudaGetLastError(); \
    if (e__ != cudaSuccess) return qb200::cuda_fail(e__, "kernel launch", __FILE__, __LINE__); } while (0)

// (re)build the separable phase tables and the projector weight list after a change of positions, lattice or species
static int nl_refresh_tables(qb200_nl* nl)
{
  if (nl->wtp_dirty) {
    nl->wtp_dirty = false;
    int rc = nl_ensure(&nl->wtp, &nl->wtp_cap, std::max(nl->Mtot, 1));
    if (rc) return rc;
    std::vector<double> w(std::max(nl->Mtot, 1), 0.0);
    for (const NlSpecies& S : nl->sp) {
      if (S.M <= 0) continue;
      std::vector<double> wt(S.npr);
      QB_CUDA(cudaMemcpy(wt.data(), S.wt, S.npr * sizeof(double), cudaMemcpyDeviceToHost));
      for (int ia = 0; ia < S.na; ia++) for (int ipr = 0; ipr < S.npr; ipr++) w[S.poff + ia * S.npr + ipr] = wt[ipr];
    }
    QB_CUDA(cudaMemcpyAsync(nl->wtp, w.data(), w.size() * sizeof(double), cudaMemcpyHostToDevice, nl->stream));
    QB_CUDA(cudaStreamSynchronize(nl->stream));
  }
  if (nl->sym_dirty) {
    nl->sym_dirty = false;
    nl->sym_ok = false;
    if (nl->gamma_half && !nl->gamma_off) {
      unsigned long long* acc = nullptr;
      QB_CUDA(cudaMalloc((void**)&acc, 2 * sizeof(unsigned long long)));
      QB_CUDA(cudaMemsetAsync(acc, 0, 2 * sizeof(unsigned long long), nl->stream));
      for (const NlSpecies& S : nl->sp) {
        if (S.M <= 0) continue;
        k_twnl_parity<<<dim3((nl->nhalf + 255) / 256, S.npr), 256, 0, nl->stream>>>(S, nl->ngw, nl->ghalf, nl->gminus, nl->nhalf, acc);
        NL_LAUNCH_CHECK(nl);
      }
      double h[2] = { 0.0, 0.0 };
      QB_CUDA(cudaMemcpyAsync(h, acc, sizeof h, cudaMemcpyDeviceToHost, nl->stream));
      QB_CUDA(cudaStreamSynchronize(nl->stream));
      cudaFree(acc);
      nl->sym_ok = h[0] <= 1e-12 * h[1];
    }
  }
  if (!nl->ph_dirty) return QB200_OK;
  nl->ph_dirty = false;
  if (!nl->lat.idx) return QB200_OK;
  for (size_t is = 0; is < nl->sp.size(); is++) {
    NlSpecies& S = nl->sp[is];
    if (S.M <= 0) continue;
    if (!nl->ph[is] || nl->ph_JT[is] < nl->lat.JT) {      // a later set_lattice may need longer tables (larger |h|,|k|,|l|)
      if (nl->ph[is]) { cudaFree(nl->ph[is]); nl->ph[is] = nullptr; }
      QB_CUDA(cudaMalloc((void**)&nl->ph[is], (size_t)S.na * nl->lat.JT * sizeof(double2)));
      nl->ph_JT[is] = nl->lat.JT;
    }
    const double* b = nl->bvec;
    k_phase_tables<<<S.na, 128, 0, nl->stream>>>(S, nl->lat, b[0], b[1], b[2], b[3], b[4], b[5], b[6], b[7], b[8],
                                                 nl->kcart[0], nl->kcart[1], nl->kcart[2], nl->ph[is]);
    NL_LAUNCH_CHECK(nl);
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
  nl->ph_dirty = true; nl->W_valid = false; nl->Wg_valid = false;
  return QB200_OK;
}

// ------------------------------------------------------------------------------------------------ update_twnl (row a11)
// NonLocalPotential::update_twnl for a Kleinman-Bylander species (NonLocalPotential.cc:261-1522, the twnl part):
// twnl[ipr][ig] = Y_lm(k+G) v(|k+G|); v = the species' radial cubic spline (Species::dvnlg, Species.cc:1492-1505: zero beyond the
// last knot of the FULL table, gcut; splintd, spline.cc:126-156), real spherical harmonics in the reference's order and
// normalisation (l=0 :334, l=1 :466-470, l=2 :712-747, l=3 :1126-1140).  One thread per plane wave, every projector of the species.
namespace qb200 {
__global__ void __launch_bounds__(128) k_twnl_kb(int ngw, const double* __restrict__ kpgx, int npr, const int* __restrict__ lproj,
                                                 const int* __restrict__ mproj, const int* __restrict__ tabproj, int nknots,
                                                 const double* __restrict__ gspl, double gcut, const double* __restrict__ ya,
                                                 const double* __restrict__ y2a, double* __restrict__ twnl)
{
  const int ig = blockIdx.x * blockDim.x + threadIdx.x;
  if (ig >= ngw) return;
  const double x = kpgx[ig], y = kpgx[ngw + ig], z = kpgx[2 * (size_t)ngw + ig];
  const double g = sqrt(x * x + y * y + z * z);
  const double gi = g > 0.0 ? 1.0 / g : 0.0;                             // Basis.cc:737
  int klo = 0, khi = nknots - 1;
  while (khi - klo > 1) { const int k = (khi + klo) >> 1; if (gspl[k] > g) khi = k; else klo = k; }
  const double h = gspl[khi] - gspl[klo], a = (gspl[khi] - g) / h, b = (g - gspl[klo]) / h;
  const double a3 = a * a * a - a, b3 = b * b * b - b, h26 = h * h * (1.0 / 6.0);
  const double pi = 3.14159265358979323846, fpi = 4.0 * pi;
  const double s14pi = sqrt(1.0 / fpi), s34pi = sqrt(3.0 / fpi), s54pi = sqrt(5.0 / fpi), s3 = sqrt(3.0), s74pi = sqrt(7.0 / fpi),
               s2132pi = sqrt(21.0 / (32. * pi)), s3532pi = sqrt(35.0 / (32. * pi)), s1054pi = sqrt(105.0 / fpi);
  const double gi2 = gi * gi, gi3 = gi2 * gi;
  const double xx = x * x * gi2, yy = y * y * gi2, zz = z * z * gi2, xy = x * y * gi2, yz = y * z * gi2, xz = x * z * gi2;
  int tlast = -1;
  double v = 0.0;
  for (int ipr = 0; ipr < npr; ipr++) {
    const int t = tabproj[ipr];
    if (t != tlast) {
      const double* Y = ya + (size_t)t * nknots;
      const double* Y2 = y2a + (size_t)t * nknots;
      v = g > gcut ? 0.0 : a * Y[klo] + b * Y[khi] + h26 * (a3 * Y2[klo] + b3 * Y2[khi]);
      tlast = t;
    }
    double ylm = 0.0;
    switch (lproj[ipr] * 8 + mproj[ipr]) {
      case 0: ylm = s14pi; break;
      case 8: ylm = s34pi * x * gi; break;
      case 9: ylm = s34pi * y * gi; break;
      case 10: ylm = s34pi * z * gi; break;
      case 16: ylm = s54pi * 0.5 * (3.0 * zz - 1.0); break;
      case 17: ylm = s54pi * 0.5 * s3 * (xx - yy); break;
      case 18: ylm = s54pi * s3 * xy; break;
      case 19: ylm = s54pi * s3 * yz; break;
      case 20: ylm = s54pi * s3 * xz; break;
      case 24: ylm = s74pi * 0.5 * z * gi * (5.0 * zz - 3.0); break;
      case 25: ylm = s2132pi * x * gi * (5.0 * zz - 1.0); break;
      case 26: ylm = s2132pi * y * gi * (5.0 * zz - 1.0); break;
      case 27: ylm = s1054pi * x * y * z * gi3; break;
      case 28: ylm = s1054pi * 0.5 * z * gi * (xx - yy); break;
      case 29: ylm = s3532pi * x * gi * (xx - 3.0 * yy); break;
      case 30: ylm = s3532pi * y * gi * (3.0 * xx - yy); break;
    }
    twnl[(size_t)ipr * ngw + ig] = ylm * v;
  }
}
}  // namespace qb200

extern "C" int qb200_nl_update_twnl(qb200_nl* nl, int is, const int* mproj, const int* tabproj, int ntab, int nknots, const double* gspl,
                                    double gcut, const double* vnlg, const double* vnlg_spl)
{
  if (!nl || is < 0 || is >= (int)nl->sp.size() || !mproj || !tabproj || ntab < 1 || nknots < 2 || !gspl || !vnlg || !vnlg_spl) {
    set_error("qb200_nl_update_twnl: bad argument"); return QB200_EINVAL;
  }
  const NlSpecies& S = nl->sp[is];
  if (S.npr == 0) return QB200_OK;
  QB_CUDA(cudaSetDevice(nl->device));
  std::vector<int> l(S.npr);
  QB_CUDA(cudaMemcpy(l.data(), S.lproj, S.npr * sizeof(int), cudaMemcpyDeviceToHost));
  for (int i = 0; i < S.npr; i++)
    if (tabproj[i] < 0 || tabproj[i] >= ntab || mproj[i] < 0 || mproj[i] > 2 * l[i]) { set_error("qb200_nl_update_twnl: projector description out of range"); return QB200_EINVAL; }
  for (int k = 1; k < nknots; k++) if (!(gspl[k] > gspl[k - 1])) { set_error("qb200_nl_update_twnl: knots must increase (spline.cc:145)"); return QB200_EINVAL; }
  void *dm = nullptr, *dt = nullptr, *dg = nullptr, *dy = nullptr, *dy2 = nullptr;
  auto fail = [&](int rc) { for (void* q : { dm, dt, dg, dy, dy2 }) if (q) cudaFree(q); return rc; };
  const size_t tb = (size_t)ntab * nknots * sizeof(double);
  if (cudaMalloc(&dm, S.npr * sizeof(int)) || cudaMalloc(&dt, S.npr * sizeof(int)) || cudaMalloc(&dg, nknots * sizeof(double)) ||
      cudaMalloc(&dy, tb) || cudaMalloc(&dy2, tb)) { set_error("qb200_nl_update_twnl: out of device memory"); return fail(QB200_ENOMEM); }
  cudaMemcpyAsync(dm, mproj, S.npr * sizeof(int), cudaMemcpyHostToDevice, nl->stream);
  cudaMemcpyAsync(dt, tabproj, S.npr * sizeof(int), cudaMemcpyHostToDevice, nl->stream);
  cudaMemcpyAsync(dg, gspl, nknots * sizeof(double), cudaMemcpyHostToDevice, nl->stream);
  cudaMemcpyAsync(dy, vnlg, tb, cudaMemcpyHostToDevice, nl->stream);
  cudaMemcpyAsync(dy2, vnlg_spl, tb, cudaMemcpyHostToDevice, nl->stream);
  k_twnl_kb<<<(nl->ngw + 127) / 128, 128, 0, nl->stream>>>(nl->ngw, nl->kpgx, S.npr, S.lproj, (const int*)dm, (const int*)dt, nknots, (const double*)dg, gcut,
                                                           (const double*)dy, (const double*)dy2, const_cast<double*>(S.twnl));
  nl->launches++;
  const cudaError_t e = cudaStreamSynchronize(nl->stream);
  fail(0);
  if (e != cudaSuccess) return cuda_fail(e, "qb200_nl_update_twnl", __FILE__, __LINE__);
  nl->W_valid = false; nl->Wg_valid = false; nl->sym_dirty = true;      // every materialised anl is stale
  return QB200_OK;
}

// Semi-local species (nquad > 0; NonLocalPotential.cc:366-419 l = 0, :500-600 l = 1, :800-960 l = 2, :1230-1345 l = 3): projector
// ipr = iquad + nquad * ilm carries twnl[ipr][ig] = Y_lm(k+G) 4 pi j_l(|k+G| r_iquad) r_iquad, the spherical Bessel functions written
// with sin / cos as the reference writes them (l = 0: 4 pi sin(q r) / q, 4 pi r at q = 0; l >= 1: 0 at q r = 0).
namespace qb200 {
__global__ void __launch_bounds__(128) k_twnl_sl(int ngw, const double* __restrict__ kpgx, int npr, const int* __restrict__ lproj,
                                                 const int* __restrict__ mproj, const double* __restrict__ rproj, double* __restrict__ twnl)
{
  const int ig = blockIdx.x * blockDim.x + threadIdx.x;
  if (ig >= ngw) return;
  const double x = kpgx[ig], y = kpgx[ngw + ig], z = kpgx[2 * (size_t)ngw + ig];
  const double g = sqrt(x * x + y * y + z * z);
  const double gi = g > 0.0 ? 1.0 / g : 0.0;
  const double pi = 3.14159265358979323846, fpi = 4.0 * pi;
  const double s14pi = sqrt(1.0 / fpi), s34pi = sqrt(3.0 / fpi), s54pi = sqrt(5.0 / fpi), s3 = sqrt(3.0), s74pi = sqrt(7.0 / fpi),
               s2132pi = sqrt(21.0 / (32. * pi)), s3532pi = sqrt(35.0 / (32. * pi)), s1054pi = sqrt(105.0 / fpi);
  const double gi2 = gi * gi, gi3 = gi2 * gi;
  const double xx = x * x * gi2, yy = y * y * gi2, zz = z * z * gi2, xy = x * y * gi2, yz = y * z * gi2, xz = x * z * gi2;
  for (int ipr = 0; ipr < npr; ipr++) {
    const int l = lproj[ipr];
    const double r = rproj[ipr], zr = g * r;
    double sn, cs;
    sincos(zr, &sn, &cs);
    double v = 0.0;
    if (l == 0) v = g == 0.0 ? fpi * r : fpi * sn * gi;
    else if (zr != 0.0) {
      const double zi = 1.0 / zr;
      if (l == 1) v = fpi * ((sn * zi - cs) * zi) * r;
      else if (l == 2) v = fpi * (((3.0 * zi * zi - 1.0) * sn - 3.0 * zi * cs) * zi) * r;
      else v = fpi * ((15.0 * zi * zi - 6.0) * zi * zi * sn - (15.0 * zi * zi - 1.0) * zi * cs) * r;
    }
    double ylm = 0.0;
    switch (l * 8 + mproj[ipr]) {
      case 0: ylm = s14pi; break;
      case 8: ylm = s34pi * x * gi; break;
      case 9: ylm = s34pi * y * gi; break;
      case 10: ylm = s34pi * z * gi; break;
      case 16: ylm = s54pi * 0.5 * (3.0 * zz - 1.0); break;
      case 17: ylm = s54pi * 0.5 * s3 * (xx - yy); break;
      case 18: ylm = s54pi * s3 * xy; break;
      case 19: ylm = s54pi * s3 * yz; break;
      case 20: ylm = s54pi * s3 * xz; break;
      case 24: ylm = s74pi * 0.5 * z * gi * (5.0 * zz - 3.0); break;
      case 25: ylm = s2132pi * x * gi * (5.0 * zz - 1.0); break;
      case 26: ylm = s2132pi * y * gi * (5.0 * zz - 1.0); break;
      case 27: ylm = s1054pi * x * y * z * gi3; break;
      case 28: ylm = s1054pi * 0.5 * z * gi * (xx - yy); break;
      case 29: ylm = s3532pi * x * gi * (xx - 3.0 * yy); break;
      case 30: ylm = s3532pi * y * gi * (3.0 * xx - yy); break;
    }
    twnl[(size_t)ipr * ngw + ig] = ylm * v;
  }
}
}  // namespace qb200

extern "C" int qb200_nl_update_twnl_semilocal(qb200_nl* nl, int is, const int* mproj, const double* rproj)
{
  if (!nl || is < 0 || is >= (int)nl->sp.size() || !mproj || !rproj) { set_error("qb200_nl_update_twnl_semilocal: bad argument"); return QB200_EINVAL; }
  const NlSpecies& S = nl->sp[is];
  if (S.npr == 0) return QB200_OK;
  QB_CUDA(cudaSetDevice(nl->device));
  std::vector<int> l(S.npr);
  QB_CUDA(cudaMemcpy(l.data(), S.lproj, S.npr * sizeof(int), cudaMemcpyDeviceToHost));
  for (int i = 0; i < S.npr; i++)
    if (mproj[i] < 0 || mproj[i] > 2 * l[i] || !(rproj[i] >= 0.0)) { set_error("qb200_nl_update_twnl_semilocal: projector description out of range"); return QB200_EINVAL; }
  void *dm = nullptr, *dr = nullptr;
  auto fail = [&](int rc) { for (void* q : { dm, dr }) if (q) cudaFree(q); return rc; };
  if (cudaMalloc(&dm, S.npr * sizeof(int)) || cudaMalloc(&dr, S.npr * sizeof(double))) { set_error("qb200_nl_update_twnl_semilocal: out of device memory"); return fail(QB200_ENOMEM); }
  cudaMemcpyAsync(dm, mproj, S.npr * sizeof(int), cudaMemcpyHostToDevice, nl->stream);
  cudaMemcpyAsync(dr, rproj, S.npr * sizeof(double), cudaMemcpyHostToDevice, nl->stream);
  k_twnl_sl<<<(nl->ngw + 127) / 128, 128, 0, nl->stream>>>(nl->ngw, nl->kpgx, S.npr, S.lproj, (const int*)dm, (const double*)dr, const_cast<double*>(S.twnl));
  nl->launches++;
  const cudaError_t e = cudaStreamSynchronize(nl->stream);
  fail(0);
  if (e != cudaSuccess) return cuda_fail(e, "qb200_nl_update_twnl_semilocal", __FILE__, __LINE__);
  nl->W_valid = false; nl->Wg_valid = false; nl->sym_dirty = true;
  return QB200_OK;
}

// the species' projector table as it sits on the device (npr * ngw doubles), for checks and for callers that keep a host copy
extern "C" int qb200_nl_get_twnl(qb200_nl* nl, int is, double* twnl)
{
  if (!nl || is < 0 || is >= (int)nl->sp.size() || !twnl) { set_error("qb200_nl_get_twnl: bad argument"); return QB200_EINVAL; }
  QB_CUDA(cudaSetDevice(nl->device));
  QB_CUDA(cudaStreamSynchronize(nl->stream));
  if (nl->sp[is].npr > 0) QB_CUDA(cudaMemcpy(twnl, nl->sp[is].twnl, (size_t)nl->sp[is].npr * nl->ngw * sizeof(double), cudaMemcpyDefault));
  return QB200_OK;
}

extern "C" int qb200_nl_set_stream(qb200_nl* nl, void* s)
{
  if (!nl) return QB200_EINVAL;
  nl->stream = (cudaStream_t)s;
  return QB200_OK;
}

extern "C" int qb200_nl_set_workspace(qb200_nl* nl, long long bytes)
{
  if (!nl || bytes < (1ll << 20)) { set_error("qb200_nl_set_workspace: bad argument"); return QB200_EINVAL; }
  nl->anl_budget = bytes;
  nl->W_valid = false; nl->Wg_valid = false;
  return QB200_OK;
}

extern "C" int qb200_nl_destroy(qb200_nl* nl)
{
  if (!nl) return QB200_OK;
  cudaSetDevice(nl->device);
  for (void* p : nl->owned) cudaFree(p);
  for (double2* p : nl->ph) if (p) cudaFree(p);
  for (void* p : nl->lat_owned) cudaFree(p);
  for (double* p : { nl->part, nl->fs, nl->eblk, nl->occ_dev, nl->enl_dev, nl->st_c, nl->st_cp, nl->wtp, nl->W, nl->Wg, nl->Ug }) if (p) cudaFree(p);
  delete nl->us;
  delete nl;
  return QB200_OK;
}

extern "C" long long qb200_nl_query(const qb200_nl* nl, int what)
{
  if (!nl) return -1;
  switch (what) {
    case 9: return nl->launches;
    case 11: return nl->nchunks_last;
    case 12: return (long long)(nl->W_cap * sizeof(double));
    case 13: return nl->Mtot;
    case 14: return nl->last_mode;
    default: return -1;
  }
}

// anl for the plane waves [gbeg, gbeg+gcount) of every species -> W (columns up to gpad zero-filled)
static int nl_generate_chunk(qb200_nl* nl, int gbeg, int gcount, int gpad, size_t WP)
{
  prof_begin(7, nl->stream);
  for (const NlSpecies& S : nl->sp) {
    if (S.M <= 0) continue;
    dim3 g((gpad + 127) / 128, S.na);
    if (nl->use3m) k_anl_gen3<<<g, 128, 0, nl->stream>>>(S, nl->lat, nl->ngw, nl->kpgx, gbeg, gcount, gpad, nl->W, WP);
    else if (nl->is_real) k_anl_gen<1><<<g, 128, 0, nl->stream>>>(S, nl->lat, nl->ngw, nl->kpgx, gbeg, gcount, gpad, nl->W, WP);
    else k_anl_gen<0><<<g, 128, 0, nl->stream>>>(S, nl->lat, nl->ngw, nl->kpgx, gbeg, gcount, gpad, nl->W, WP);
    NL_LAUNCH_CHECK(nl);
  }
  prof_end(nl->stream);
  return QB200_OK;
}

// plane-wave chunks the projector sweep of a call takes (1: anl for the whole sphere fits the workspace)
static void nl_chunking(const qb200_nl* nl, int* gchunk_out, int* nchunks_out)
{
  // bytes of W per plane wave: Gamma Mtot rows x 2 doubles; complex 2*Mtot rows x 2 doubles, or (3M) 24*ceil(Mtot/8) x 1
  const long long per_g = nl->is_real ? 16ll * nl->Mtot : (nl->use3m ? 8ll * 192 * ((nl->Mtot + 63) / 64) : 32ll * nl->Mtot);
  long long gmax = nl->anl_budget / std::max(per_g, 1ll);
  gmax = std::max(512ll, (gmax / 512) * 512);
  const int gchunk = (int)std::min<long long>(gmax, ((long long)nl->ngw + 15) / 16 * 16);
  *gchunk_out = gchunk;
  *nchunks_out = (nl->ngw + gchunk - 1) / gchunk;
}
int qb200_nl_projectors(const qb200_nl* nl) { return nl->Mtot; }
int qb200_nl_chunks(const qb200_nl* nl, int) { int g, n; nl_chunking(nl, &g, &n); return nl->Mtot > 0 ? n : 1; }

// Gamma point, complex states: the projector contraction over the half sphere (kernels above).  occ already on the device,
// enl_dev cleared by the caller, tables refreshed.
static int nl_energy_gamma_half(qb200_nl* nl, int ldc, int nst, const double* c, int compute_hpsi, double* cp, int cont, int overwrite)
{
  int rc;
  nl->last_mode = 3;
  nl->nchunks_last = 1;
  const int Mtot = nl->Mtot, nhalf = nl->nhalf, ngw = nl->ngw;
  const int hpad = (nhalf + 15) / 16 * 16;
  const size_t WP = 2 * (size_t)hpad, ldu = (size_t)hpad;
  const int Mp = (Mtot + 1) & ~1;
  const int nst2 = 2 * nst;
  if (nl->Wg_cap < (size_t)Mtot * WP) nl->Wg_valid = false;
  if ((rc = nl_ensure(&nl->Wg, &nl->Wg_cap, (size_t)Mtot * WP))) return rc;
  if ((rc = nl_ensure(&nl->Ug, &nl->Ug_cap, 2 * ldu * nst2))) return rc;
  if (!(nl->Wg_valid && (nl->cache_anl || cont))) {
    prof_begin(7, nl->stream);
    for (const NlSpecies& S : nl->sp) {
      if (S.M <= 0) continue;
      k_anl_gen_half<<<dim3((hpad + 127) / 128, S.na), 128, 0, nl->stream>>>(S, nl->lat, ngw, nl->kpgx, nl->ghalf, nhalf, hpad, nl->Wg, WP);
      NL_LAUNCH_CHECK(nl);
    }
    prof_end(nl->stream);
    nl->Wg_valid = true;
  }
  // split-K of k_fnl<1>: 128 x 128 tiles, one CTA per SM
  // equal column tiles: 2*nst = 192 real columns run as 2 x 96, not 128 + 64 (QB200_NL_TN=0: full tiles and a remainder)
  int nt = (nst2 + NL_TN - 1) / NL_TN, tn = NL_TN;
  { const char* e = getenv("QB200_NL_TN"); if (!(e && atoi(e) == 0)) { tn = std::min(NL_TN, ((nst2 + nt - 1) / nt + 31) / 32 * 32); nt = (nst2 + tn - 1) / tn; } }
  const int mt = (Mtot + NL_TM - 1) / NL_TM;
  int ksplit = 1;
  if (const char* e = getenv("QB200_NL_KSPLIT")) ksplit = std::max(1, std::min(64, atoi(e)));
  else {
    const int maxk = std::max(1, nhalf / 256);
    double best = -1.0;
    for (int k = 1; k <= std::min(maxk, 64); k++) {
      const long ctas = (long)mt * nt * k;
      const long waves = (ctas + nl->nsm - 1) / nl->nsm;
      const double eff = (double)ctas / (double)(waves * nl->nsm) - 0.002 * k;
      if (eff > best) { best = eff; ksplit = k; }
    }
  }
  if ((rc = nl_ensure(&nl->part, &nl->part_cap, (size_t)ksplit * nst2 * Mp))) return rc;
  if ((rc = nl_ensure(&nl->fs, &nl->fs_cap, (size_t)nst2 * Mp))) return rc;
  if (Mp != Mtot) QB_CUDA(cudaMemsetAsync(nl->fs, 0, (size_t)nst2 * Mp * sizeof(double), nl->stream));
  const size_t total = (size_t)nst * Mtot;
  const int nblk = (int)((total + 255) / 256);
  if ((rc = nl_ensure(&nl->eblk, &nl->eblk_cap, nblk))) return rc;
  const dim3 gpm((nhalf + 255) / 256, nst);
  prof_begin(3, nl->stream);
  k_split_pm<<<gpm, 256, 0, nl->stream>>>((const double2*)c, ldc, nl->ghalf, nl->gminus, nhalf, (double2*)nl->Ug, ldu);
  NL_LAUNCH_CHECK(nl);
  int kper = (2 * nhalf + ksplit - 1) / ksplit;
  kper = (kper + NL_KSTEP - 1) / NL_KSTEP * NL_KSTEP;
  k_fnl<1><<<dim3(mt, nt, ksplit), NL_THREADS, FNL_SMEM_BYTES, nl->stream>>>(nl->Wg, WP, Mtot, 0, nhalf, kper, (const double2*)nl->Ug, ldu, nst2, nl->part, Mp, Mtot, 0, tn);
  prof_end(nl->stream);
  NL_LAUNCH_CHECK(nl);
  prof_begin(4, nl->stream);
  k_fnl_finish_half<<<nblk, 256, 0, nl->stream>>>(nl->wtp, Mtot, nl->part, Mp, nst, ksplit, nl->occ_dev, 1.0 / nl->omega, nl->fs, nl->eblk);
  NL_LAUNCH_CHECK(nl);
  k_sum_blocks<<<1, 256, 0, nl->stream>>>(nl->eblk, nblk, nl->enl_dev);
  prof_end(nl->stream);
  NL_LAUNCH_CHECK(nl);
  if (!compute_hpsi) return QB200_OK;
  prof_begin(5, nl->stream);
  k_back<2><<<dim3(nt, (nhalf + 63) / 64), NL_THREADS, BK_SMEM_BYTES, nl->stream>>>(nl->Wg, WP, Mtot, 0, nhalf, nl->fs, Mp, (double2*)cp, ldc, nst2, overwrite,
                                                                                    nl->ghalf, nl->gminus, tn);
  prof_end(nl->stream);
  NL_LAUNCH_CHECK(nl);
  return QB200_OK;
}

// device pointers; enl accumulated into nl->enl_dev.  flags bit 0 clear: a new call (enl zeroed, anl regenerated unless
// cached); bit 0 set: a further block of states of the same call (enl keeps accumulating, a whole-sphere anl in W is
// reused).  Bit 1: cp rows [0, ngw) are known to be zero and are WRITTEN instead of accumulated (H psi: first term).
int qb200_nl_energy_dev(qb200_nl* nl, int ldc, int nst, const double* c, const double* occ_host, int compute_hpsi, double* cp, int flags)
{
  const int cont = flags & 1, overwrite = (flags >> 1) & 1;
  int rc;
  if ((rc = nl_ensure(&nl->occ_dev, &nl->occ_cap, nst))) return rc;
  QB_CUDA(cudaMemcpyAsync(nl->occ_dev, occ_host, nst * sizeof(double), cudaMemcpyDefault, nl->stream));
  if (!cont) QB_CUDA(cudaMemsetAsync(nl->enl_dev, 0, sizeof(double), nl->stream));
  const int Mtot = nl->Mtot;
  if (Mtot <= 0) return QB200_OK;
  if ((rc = nl_refresh_tables(nl))) return rc;
  if (nl->gamma_half && !nl->gamma_off && nl->sym_ok) {
    // whole half-sphere anl + the two half-sphere blocks must fit the workspace; otherwise the chunked sweeps below
    const size_t hpad = ((size_t)nl->nhalf + 15) / 16 * 16;
    const long long need = 8ll * (long long)Mtot * 2 * (long long)hpad;
    if (need <= nl->anl_budget) return nl_energy_gamma_half(nl, ldc, nst, c, compute_hpsi, cp, cont, overwrite);
  }
  const int real = nl->is_real;
  nl->last_mode = real ? 0 : (nl->use3m ? 2 : 1);
  const int ncols = real ? nst : 2 * nst;
  const bool m3 = nl->use3m && !real;              // Karatsuba form (nonlocal_3m.cuh)
  const int RW = real ? Mtot : (m3 ? 24 * ((Mtot + 7) / 8) : 2 * Mtot);          // rows of W
  const int Mp = (Mtot + 1) & ~1;                  // even pitch: 16-byte copies of fs stay aligned; the pad is zero
  const int FP = real ? Mp : (m3 ? RW : 2 * Mp);
  // chunking of the plane waves: W = RW x 2*gchunk doubles within the budget
  const int ngw = nl->ngw;
  int gchunk, nchunks;
  nl_chunking(nl, &gchunk, &nchunks);
  nl->nchunks_last = nchunks;
  const size_t WP = m3 ? (size_t)gchunk : 2 * (size_t)gchunk;
  // 3M: W3 is allocated for whole k_fnl3 tiles (64 projectors) plus one k_back3 tile of slack past the last row; rows
  // that hold no projector stay zero from the fill below
  const size_t Welems = m3 ? (size_t)192 * ((Mtot + 63) / 64) * WP + 128 : (size_t)RW * WP;
  if (nl->W_cap < Welems) { nl->W_valid = false; nl->W_WP = 0; }
  if ((rc = nl_ensure(&nl->W, &nl->W_cap, Welems))) return rc;
  if (m3 && nl->W_WP != WP) {
    QB_CUDA(cudaMemsetAsync(nl->W, 0, nl->W_cap * sizeof(double), nl->stream));
    nl->W_WP = WP; nl->W_valid = false;
  }
  // W_valid: W holds anl of the whole sphere for the current positions; reused by later blocks of one call, and
  // across calls when caching is on
  const bool reuse = nchunks == 1 && nl->W_valid && (nl->cache_anl || cont);
  // split K of k_fnl so that the CTAs fill whole waves of the SMs (one CTA per SM)
  const int nt3 = nl->tile3m == 1 ? 64 : 128;      // states per CTA tile of the 3M kernels
  const int mt = m3 ? (Mtot + 63) / 64 : (RW + NL_TM - 1) / NL_TM, nt = m3 ? (nst + nt3 - 1) / nt3 : (nst + NL_TN - 1) / NL_TN;
  const int slots = (m3 && nl->tile3m == 1) ? 2 * nl->nsm : nl->nsm;   // resident CTAs
  int ksplit = 1;
  {
    const int maxk = std::max(1, std::min(gchunk, ngw) / 512);
    double best = -1.0;
    for (int k = 1; k <= std::min(maxk, 64); k++) {
      const long ctas = (long)mt * nt * k;
      const long waves = (ctas + slots - 1) / slots;
      const double eff = (double)ctas / (double)(waves * slots) - 0.002 * k;   // mild preference for fewer partials
      if (eff > best) { best = eff; ksplit = k; }
    }
  }
  if ((rc = nl_ensure(&nl->part, &nl->part_cap, (size_t)ksplit * ncols * Mp))) return rc;
  if ((rc = nl_ensure(&nl->fs, &nl->fs_cap, (size_t)nst * FP))) return rc;
  if (Mp != Mtot || (m3 && RW != 3 * Mtot)) QB_CUDA(cudaMemsetAsync(nl->fs, 0, (size_t)nst * FP * sizeof(double), nl->stream));
  const size_t total = (size_t)nst * Mtot;
  const int nblk = (int)((total + 255) / 256);
  if ((rc = nl_ensure(&nl->eblk, &nl->eblk_cap, nblk))) return rc;
  // sweep 1: fnl partials, chunk by chunk
  for (int ch = 0; ch < nchunks; ch++) {
    const int gbeg = ch * gchunk, gcount = std::min(gchunk, ngw - gbeg), gpad = (gcount + 15) / 16 * 16;
    if (!reuse) {
      if ((rc = nl_generate_chunk(nl, gbeg, gcount, gpad, WP))) return rc;
      nl->W_valid = nchunks == 1;
    }
    int kper = ((m3 ? 1 : 2) * gcount + ksplit - 1) / ksplit;      // reduction index: plane waves (3M) or their reals
    kper = m3 ? (kper + N3_KS - 1) / N3_KS * N3_KS : (kper + NL_KSTEP - 1) / NL_KSTEP * NL_KSTEP;
    dim3 g1(mt, nt, ksplit);          // a split beyond the chunk's end has no stages and stores zeros
    prof_begin(3, nl->stream);
    if (m3 && nl->tile3m == 1) k_fnl3<4, 2, 2><<<g1, 256, Fnl3Cfg<4, 2, 2>::SMEM, nl->stream>>>(nl->W, WP, gbeg, gcount, kper, (const double2*)c, ldc, nst, nl->part, Mp, Mtot, ch > 0);
    else if (m3) k_fnl3<4, 4, 3><<<g1, 512, Fnl3Cfg<4, 4, 3>::SMEM, nl->stream>>>(nl->W, WP, gbeg, gcount, kper, (const double2*)c, ldc, nst, nl->part, Mp, Mtot, ch > 0);
    else if (real) k_fnl<1><<<g1, NL_THREADS, FNL_SMEM_BYTES, nl->stream>>>(nl->W, WP, RW, gbeg, gcount, kper, (const double2*)c, ldc, nst, nl->part, Mp, Mtot, ch > 0);
    else k_fnl<0><<<g1, NL_THREADS, FNL_SMEM_BYTES, nl->stream>>>(nl->W, WP, RW, gbeg, gcount, kper, (const double2*)c, ldc, nst, nl->part, Mp, Mtot, ch > 0);
    prof_end(nl->stream);
    NL_LAUNCH_CHECK(nl);
  }
  prof_begin(4, nl->stream);
  if (m3) k_fnl_finish3<<<nblk, 256, 0, nl->stream>>>(nl->wtp, Mtot, nl->part, Mp, nst, ksplit, nl->occ_dev, 1.0 / nl->omega, nl->fs, FP, nl->eblk);
  else if (real) k_fnl_finish<1><<<nblk, 256, 0, nl->stream>>>(nl->wtp, Mtot, nl->part, Mp, nst, ksplit, nl->occ_dev, 1.0 / nl->omega, nl->fs, nl->eblk);
  else k_fnl_finish<0><<<nblk, 256, 0, nl->stream>>>(nl->wtp, Mtot, nl->part, Mp, nst, ksplit, nl->occ_dev, 1.0 / nl->omega, nl->fs, nl->eblk);
  NL_LAUNCH_CHECK(nl);
  k_sum_blocks<<<1, 256, 0, nl->stream>>>(nl->eblk, nblk, nl->enl_dev);
  prof_end(nl->stream);
  NL_LAUNCH_CHECK(nl);
  if (!compute_hpsi) return QB200_OK;
  // sweep 2: back-projection, chunk by chunk (the last chunk of sweep 1 is still in W)
  for (int i = 0; i < nchunks; i++) {
    const int ch = nchunks - 1 - i;
    const int gbeg = ch * gchunk, gcount = std::min(gchunk, ngw - gbeg), gpad = (gcount + 15) / 16 * 16;
    if (i > 0 && (rc = nl_generate_chunk(nl, gbeg, gcount, gpad, WP))) return rc;   // (i == 0: still in W from sweep 1)
    dim3 g2(nt, (gcount + 63) / 64);
    prof_begin(5, nl->stream);
    if (m3 && nl->tile3m == 1) k_back3<4, 2, 3><<<dim3(nt, (gcount + 63) / 64), 256, Back3Cfg<4, 2, 3>::SMEM, nl->stream>>>(nl->W, WP, RW, gbeg, gcount, nl->fs, FP, (double2*)cp, ldc, nst, overwrite);
    else if (m3) k_back3<4, 4, 4><<<dim3(nt, (gcount + 63) / 64), 512, Back3Cfg<4, 4, 4>::SMEM, nl->stream>>>(nl->W, WP, RW, gbeg, gcount, nl->fs, FP, (double2*)cp, ldc, nst, overwrite);
    else if (real) k_back<1><<<g2, NL_THREADS, BK_SMEM_BYTES, nl->stream>>>(nl->W, WP, RW, gbeg, gcount, nl->fs, FP, (double2*)cp, ldc, nst, overwrite);
    else k_back<0><<<g2, NL_THREADS, BK_SMEM_BYTES, nl->stream>>>(nl->W, WP, RW, gbeg, gcount, nl->fs, FP, (double2*)cp, ldc, nst, overwrite);
    prof_end(nl->stream);
    NL_LAUNCH_CHECK(nl);
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
  if ((rc = qb200_nl_energy_dev(nl, ldc, nst, cd, occ, compute_hpsi, cpd, 0))) return rc;
  if (compute_hpsi && cpd != cp) QB_CUDA(cudaMemcpyAsync(cp, cpd, blk * sizeof(double), cudaMemcpyDeviceToHost, nl->stream));
  double e = 0.0;
  QB_CUDA(cudaMemcpyAsync(&e, nl->enl_dev, sizeof(double), cudaMemcpyDeviceToHost, nl->stream));
  QB_CUDA(cudaStreamSynchronize(nl->stream));
  if (enl) *enl = e;
  return QB200_OK;
}

// ------------------------------------------------------------------------------------------------ ultrasoft beta.psi path
// SlaterDet::calc_betapsi (SlaterDet.cc:2130-2263) and the gemms of SlaterDet::calc_spsi (:2426-2570)
// are the projector contraction and the back-projection WITHOUT the diagonal weights of the norm-conserving form: betapsi =
// anl^H c, S psi = psi + anl (q betapsi) / omega with the species' symmetric coupling q between the channels of one atom.
// The tables are the reference's betag (beta_b(|k+G|) Y_lm(k+G), SlaterDet::calc_betag :2006-2127; an INPUT like twnl) given
// to qb200_nl_add_species with lproj = l of the channel; the structure factor and (-i)^l are applied as for anl.  Ultrasoft
// potentials force complex states (SlaterDet.cc:57-58), so these entry points require a complex basis; they run the
// 4-product DMMA kernels (k_anl_gen<0>, k_fnl<0>, k_back<0>), plane-wave chunk by chunk.
__global__ void __launch_bounds__(256) k_us_collect(const double* __restrict__ part, int Mp, int nst, int ksplit, int Mtot, double* __restrict__ bp)
{
  const size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (idx >= (size_t)nst * Mtot) return;
  const int n = (int)(idx / Mtot), p = (int)(idx % Mtot);
  double fr = 0.0, fi = 0.0;
  for (int ks = 0; ks < ksplit; ks++) {                     // fixed order: deterministic
    fr += part[((size_t)ks * 2 * nst + 2 * n) * Mp + p];
    fi += part[((size_t)ks * 2 * nst + 2 * n + 1) * Mp + p];
  }
  bp[2 * idx] = fr;
  bp[2 * idx + 1] = fi;
}
// fs[n][p] (pitch Mp complex, zero pad) = f[n][p]
__global__ void __launch_bounds__(256) k_us_pack(const double2* __restrict__ f, int Mtot, int nst, int Mp, double2* __restrict__ fs)
{
  const size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (idx >= (size_t)nst * Mp) return;
  const int n = (int)(idx / Mp), p = (int)(idx % Mp);
  fs[idx] = p < Mtot ? f[(size_t)n * Mtot + p] : make_double2(0.0, 0.0);
}
// f[n][p] = omega_inv * sum_lm' q_species(p)[lm(p)][lm'] bp[n][atom block of p + lm'];  meta[4p..] = block start, npr, q offset, lm
__global__ void __launch_bounds__(256) k_us_couple(const double2* __restrict__ bp, int Mtot, int nst, const int* __restrict__ meta,
                                                   const double* __restrict__ q, double omega_inv, double2* __restrict__ f)
{
  const size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (idx >= (size_t)nst * Mtot) return;
  const int n = (int)(idx / Mtot), p = (int)(idx % Mtot);
  const int b0 = meta[4 * p], npr = meta[4 * p + 1], qo = meta[4 * p + 2], lm = meta[4 * p + 3];
  const double2* b = bp + (size_t)n * Mtot + b0;
  const double* qr = q + qo + (size_t)lm * npr;
  double fr = 0.0, fi = 0.0;
  for (int j = 0; j < npr; j++) { fr += qr[j] * b[j].x; fi += qr[j] * b[j].y; }
  f[idx] = make_double2(omega_inv * fr, omega_inv * fi);
}

static int nl_us_check(qb200_nl* nl, const char* who)
{
  if (nl->is_real) { set_error(std::string(who) + ": ultrasoft projectors need a complex basis (SlaterDet.cc:57-58)"); return QB200_EUNSUPPORTED; }
  return QB200_OK;
}
static void nl_us_chunking(const qb200_nl* nl, int* gchunk, int* nchunks)
{
  long long gmax = nl->anl_budget / std::max(32ll * nl->Mtot, 1ll);
  gmax = std::max(512ll, (gmax / 512) * 512);
  *gchunk = (int)std::min<long long>(gmax, ((long long)nl->ngw + 15) / 16 * 16);
  *nchunks = (nl->ngw + *gchunk - 1) / *gchunk;
}
static int nl_us_gen(qb200_nl* nl, int gbeg, int gcount, int gpad, size_t WP)
{
  for (const NlSpecies& S : nl->sp) {
    if (S.M <= 0) continue;
    k_anl_gen<0><<<dim3((gpad + 127) / 128, S.na), 128, 0, nl->stream>>>(S, nl->lat, nl->ngw, nl->kpgx, gbeg, gcount, gpad, nl->W, WP);
    NL_LAUNCH_CHECK(nl);
  }
  return QB200_OK;
}
// bp_dev[nst][Mtot] complex (device) = anl^H c
static int nl_us_project(qb200_nl* nl, int ldc, int nst, const double* c, double* bp_dev)
{
  int rc;
  if ((rc = nl_refresh_tables(nl))) return rc;
  const int Mtot = nl->Mtot, RW = 2 * Mtot, Mp = (Mtot + 1) & ~1, ngw = nl->ngw;
  int gchunk, nchunks;
  nl_us_chunking(nl, &gchunk, &nchunks);
  const size_t WP = 2 * (size_t)gchunk;
  if ((rc = nl_ensure(&nl->W, &nl->W_cap, (size_t)RW * WP))) return rc;
  nl->W_valid = false; nl->W_WP = 0;                                     // W is rewritten in the 4-product layout
  const int mt = (RW + NL_TM - 1) / NL_TM, nt = (nst + NL_TN - 1) / NL_TN;
  int ksplit = 1;
  {
    const int maxk = std::max(1, std::min(gchunk, ngw) / 512);
    double best = -1.0;
    for (int k = 1; k <= std::min(maxk, 64); k++) {
      const long ctas = (long)mt * nt * k, waves = (ctas + nl->nsm - 1) / nl->nsm;
      const double eff = (double)ctas / (double)(waves * nl->nsm) - 0.002 * k;
      if (eff > best) { best = eff; ksplit = k; }
    }
  }
  if ((rc = nl_ensure(&nl->part, &nl->part_cap, (size_t)ksplit * 2 * nst * Mp))) return rc;
  for (int ch = 0; ch < nchunks; ch++) {
    const int gbeg = ch * gchunk, gcount = std::min(gchunk, ngw - gbeg), gpad = (gcount + 15) / 16 * 16;
    if ((rc = nl_us_gen(nl, gbeg, gcount, gpad, WP))) return rc;
    int kper = (2 * gcount + ksplit - 1) / ksplit;
    kper = (kper + NL_KSTEP - 1) / NL_KSTEP * NL_KSTEP;
    k_fnl<0><<<dim3(mt, nt, ksplit), NL_THREADS, FNL_SMEM_BYTES, nl->stream>>>(nl->W, WP, RW, gbeg, gcount, kper, (const double2*)c, ldc, nst, nl->part, Mp, Mtot, ch > 0);
    NL_LAUNCH_CHECK(nl);
  }
  const size_t total = (size_t)nst * Mtot;
  k_us_collect<<<(unsigned)((total + 255) / 256), 256, 0, nl->stream>>>(nl->part, Mp, nst, ksplit, Mtot, bp_dev);
  NL_LAUNCH_CHECK(nl);
  return QB200_OK;
}
// cp (device) += anl f, f_dev[nst][Mtot] complex (device)
static int nl_us_backproject(qb200_nl* nl, int ldc, int nst, const double* f_dev, double* cp)
{
  int rc;
  if ((rc = nl_refresh_tables(nl))) return rc;
  const int Mtot = nl->Mtot, RW = 2 * Mtot, Mp = (Mtot + 1) & ~1, ngw = nl->ngw;
  int gchunk, nchunks;
  nl_us_chunking(nl, &gchunk, &nchunks);
  const size_t WP = 2 * (size_t)gchunk;
  if ((rc = nl_ensure(&nl->W, &nl->W_cap, (size_t)RW * WP))) return rc;
  nl->W_valid = false; nl->W_WP = 0;
  if ((rc = nl_ensure(&nl->fs, &nl->fs_cap, (size_t)nst * 2 * Mp))) return rc;
  k_us_pack<<<(unsigned)(((size_t)nst * Mp + 255) / 256), 256, 0, nl->stream>>>((const double2*)f_dev, Mtot, nst, Mp, (double2*)nl->fs);
  NL_LAUNCH_CHECK(nl);
  const int nt = (nst + NL_TN - 1) / NL_TN;
  for (int ch = 0; ch < nchunks; ch++) {
    const int gbeg = ch * gchunk, gcount = std::min(gchunk, ngw - gbeg), gpad = (gcount + 15) / 16 * 16;
    if ((rc = nl_us_gen(nl, gbeg, gcount, gpad, WP))) return rc;
    k_back<0><<<dim3(nt, (gcount + 63) / 64), NL_THREADS, BK_SMEM_BYTES, nl->stream>>>(nl->W, WP, RW, gbeg, gcount, nl->fs, 2 * Mp, (double2*)cp, ldc, nst, 0);
    NL_LAUNCH_CHECK(nl);
  }
  return QB200_OK;
}

// small device scratch of the ultrasoft entry points (freed with the object through `owned`)
static int nl_us_scratch(qb200_nl* nl, size_t bytes, void** out)
{
  void* d = nullptr;
  QB_CUDA(cudaMalloc(&d, std::max<size_t>(bytes, 16)));
  *out = d;
  return QB200_OK;
}

extern "C" int qb200_nl_betapsi(qb200_nl* nl, int ldc, int nst, const double* c, double* betapsi)
{
  if (!nl || !c || !betapsi || nst < 0 || ldc < nl->ngw) { set_error("qb200_nl_betapsi: bad argument"); return QB200_EINVAL; }
  int rc;
  if ((rc = nl_us_check(nl, "qb200_nl_betapsi"))) return rc;
  if (nst == 0 || nl->Mtot == 0) return QB200_OK;
  QB_CUDA(cudaSetDevice(nl->device));
  const size_t blk = 2 * (size_t)ldc * nst, nbp = 2 * (size_t)nst * nl->Mtot;
  const double* cd = c;
  if (!is_device_ptr(c)) {
    if ((rc = nl_ensure(&nl->st_c, &nl->st_c_cap, blk))) return rc;
    QB_CUDA(cudaMemcpyAsync(nl->st_c, c, blk * sizeof(double), cudaMemcpyHostToDevice, nl->stream));
    cd = nl->st_c;
  }
  void* bpd = betapsi;
  const bool bhost = !is_device_ptr(betapsi);
  if (bhost && (rc = nl_us_scratch(nl, nbp * sizeof(double), &bpd))) return rc;
  rc = nl_us_project(nl, ldc, nst, cd, (double*)bpd);
  if (!rc && bhost) { const cudaError_t e = cudaMemcpyAsync(betapsi, bpd, nbp * sizeof(double), cudaMemcpyDeviceToHost, nl->stream); if (e != cudaSuccess) rc = cuda_fail(e, "copy", __FILE__, __LINE__); }
  const cudaError_t es = cudaStreamSynchronize(nl->stream);
  if (bhost) cudaFree(bpd);
  if (!rc && es != cudaSuccess) rc = cuda_fail(es, "qb200_nl_betapsi", __FILE__, __LINE__);
  return rc;
}

extern "C" int qb200_nl_add_beta(qb200_nl* nl, int ldc, int nst, const double* f, double* cp)
{
  if (!nl || !f || !cp || nst < 0 || ldc < nl->ngw) { set_error("qb200_nl_add_beta: bad argument"); return QB200_EINVAL; }
  int rc;
  if ((rc = nl_us_check(nl, "qb200_nl_add_beta"))) return rc;
  if (nst == 0 || nl->Mtot == 0) return QB200_OK;
  QB_CUDA(cudaSetDevice(nl->device));
  const size_t blk = 2 * (size_t)ldc * nst, nbp = 2 * (size_t)nst * nl->Mtot;
  void* fd = const_cast<double*>(f);
  const bool fhost = !is_device_ptr(f);
  if (fhost) {
    if ((rc = nl_us_scratch(nl, nbp * sizeof(double), &fd))) return rc;
    QB_CUDA(cudaMemcpyAsync(fd, f, nbp * sizeof(double), cudaMemcpyHostToDevice, nl->stream));
  }
  double* cpd = cp;
  if (!is_device_ptr(cp)) {
    if ((rc = nl_ensure(&nl->st_cp, &nl->st_cp_cap, blk))) return rc;
    QB_CUDA(cudaMemcpyAsync(nl->st_cp, cp, blk * sizeof(double), cudaMemcpyHostToDevice, nl->stream));
    cpd = nl->st_cp;
  }
  rc = nl_us_backproject(nl, ldc, nst, (const double*)fd, cpd);
  if (!rc && cpd != cp) { const cudaError_t e = cudaMemcpyAsync(cp, cpd, blk * sizeof(double), cudaMemcpyDeviceToHost, nl->stream); if (e != cudaSuccess) rc = cuda_fail(e, "copy", __FILE__, __LINE__); }
  const cudaError_t es = cudaStreamSynchronize(nl->stream);
  if (fhost) cudaFree(fd);
  if (!rc && es != cudaSuccess) rc = cuda_fail(es, "qb200_nl_add_beta", __FILE__, __LINE__);
  return rc;
}

extern "C" int qb200_nl_spsi(qb200_nl* nl, int ldc, int nst, const double* c, const double* qmat, double* spsi, double* betapsi)
{
  if (!nl || !c || !qmat || !spsi || nst < 0 || ldc < nl->ngw) { set_error("qb200_nl_spsi: bad argument"); return QB200_EINVAL; }
  int rc;
  if ((rc = nl_us_check(nl, "qb200_nl_spsi"))) return rc;
  if (nst == 0) return QB200_OK;
  QB_CUDA(cudaSetDevice(nl->device));
  const size_t blk = 2 * (size_t)ldc * nst, nbp = 2 * (size_t)nst * nl->Mtot;
  const double* cd = c;
  if (!is_device_ptr(c)) {
    if ((rc = nl_ensure(&nl->st_c, &nl->st_c_cap, blk))) return rc;
    QB_CUDA(cudaMemcpyAsync(nl->st_c, c, blk * sizeof(double), cudaMemcpyHostToDevice, nl->stream));
    cd = nl->st_c;
  }
  double* sd = spsi;
  if (!is_device_ptr(spsi)) {
    if ((rc = nl_ensure(&nl->st_cp, &nl->st_cp_cap, blk))) return rc;
    sd = nl->st_cp;
  }
  QB_CUDA(cudaMemcpyAsync(sd, cd, blk * sizeof(double), cudaMemcpyDeviceToDevice, nl->stream));   // spsi_ = c_  (SlaterDet.cc:2438)
  if (nl->Mtot > 0) {
    // per-projector description of the coupling and the species' q matrices
    std::vector<int> meta(4 * (size_t)nl->Mtot);
    size_t qo = 0;
    for (const NlSpecies& S : nl->sp) {
      for (int ia = 0; ia < S.na; ia++)
        for (int lm = 0; lm < S.npr; lm++) {
          const size_t p = (size_t)S.poff + (size_t)ia * S.npr + lm;
          meta[4 * p] = S.poff + ia * S.npr; meta[4 * p + 1] = S.npr; meta[4 * p + 2] = (int)qo; meta[4 * p + 3] = lm;
        }
      qo += (size_t)S.npr * S.npr;
    }
    void *md = nullptr, *qd = nullptr, *bpd = nullptr, *fd = nullptr;
    if ((rc = nl_us_scratch(nl, meta.size() * sizeof(int), &md)) || (rc = nl_us_scratch(nl, qo * sizeof(double), &qd)) ||
        (rc = nl_us_scratch(nl, nbp * sizeof(double), &bpd)) || (rc = nl_us_scratch(nl, nbp * sizeof(double), &fd))) {
      for (void* q : { md, qd, bpd, fd }) if (q) cudaFree(q);
      return rc;
    }
    cudaMemcpyAsync(md, meta.data(), meta.size() * sizeof(int), cudaMemcpyHostToDevice, nl->stream);
    cudaMemcpyAsync(qd, qmat, qo * sizeof(double), cudaMemcpyDefault, nl->stream);
    rc = nl_us_project(nl, ldc, nst, cd, (double*)bpd);
    if (!rc) {
      const size_t total = (size_t)nst * nl->Mtot;
      k_us_couple<<<(unsigned)((total + 255) / 256), 256, 0, nl->stream>>>((const double2*)bpd, nl->Mtot, nst, (const int*)md, (const double*)qd, 1.0 / nl->omega, (double2*)fd);
      nl->launches++;
      rc = nl_us_backproject(nl, ldc, nst, (const double*)fd, sd);
    }
    if (!rc && betapsi) cudaMemcpyAsync(betapsi, bpd, nbp * sizeof(double), cudaMemcpyDefault, nl->stream);
    const cudaError_t es = cudaStreamSynchronize(nl->stream);
    for (void* q : { md, qd, bpd, fd }) cudaFree(q);
    if (!rc && es != cudaSuccess) rc = cuda_fail(es, "qb200_nl_spsi", __FILE__, __LINE__);
    if (rc) return rc;
  }
  if (sd != spsi) QB_CUDA(cudaMemcpyAsync(spsi, sd, blk * sizeof(double), cudaMemcpyDeviceToHost, nl->stream));
  QB_CUDA(cudaStreamSynchronize(nl->stream));
  return QB200_OK;
}

#include "ultrasoft.cuh"

double* qb200_nl_enl_dev(qb200_nl* nl) { return nl->enl_dev; }

// E_nl of the last qb200_nl_energy / qb200_hpsi call (the device scalar the kernels summed into).  A DEVICE destination is
// written on the object's stream without synchronising -- for callers that passed enl = NULL and reduce the energy on the
// device, e.g. appended to the density buffer so that ONE all-reduce carries rho and E_nl; a HOST destination synchronises.
extern "C" int qb200_nl_last_enl(qb200_nl* nl, double* enl)
{
  if (!nl || !enl) { set_error("qb200_nl_last_enl: bad argument"); return QB200_EINVAL; }
  QB_CUDA(cudaSetDevice(nl->device));
  if (is_device_ptr(enl)) {
    QB_CUDA(cudaMemcpyAsync(enl, nl->enl_dev, sizeof(double), cudaMemcpyDeviceToDevice, nl->stream));
  } else {
    QB_CUDA(cudaMemcpyAsync(enl, nl->enl_dev, sizeof(double), cudaMemcpyDeviceToHost, nl->stream));
    QB_CUDA(cudaStreamSynchronize(nl->stream));
  }
  return QB200_OK;
}
cudaStream_t qb200_nl_swap_stream(qb200_nl* nl, cudaStream_t s) { cudaStream_t o = nl->stream; nl->stream = s; return o; }

// the subspace dense linear algebra on the same GEMM kernels (qb200_residual, qb200_gram)
#include "subspace_la.cuh"
