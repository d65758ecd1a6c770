// qball_b200/csrc/nonlocal_3m.cuh -- the projector contractions of complex bases as THREE real DMMA GEMMs each
// (Karatsuba / "3M" complex product) instead of the four of the straightforward real embedding: 25 % fewer FP64
// tensor flops for the path's one FP64-pipe-bound stage (NonLocalPotential.cc:2050-2068 fnl = anl^H c and
// :2150-2171 cp += anl (wt/omega fnl), both zgemm in the reference).
//
// With anl = A + iB (A, B real), c = X + iY, fs = F + iH:
//   fnl = anl^* . c :  P1 = sum A X,  P2 = sum B Y,  P3 = sum (A+B)(Y-X)   ->  Re = P1 + P2,  Im = P3 + P1 - P2
//   cp += anl . fs  :  R1 = sum A F,  R2 = sum B H,  R3 = sum (A+B)(F+H)   ->  Re = R1 - R2,  Im = R3 - R1 - R2
// The three "kinds" of the anl operand (A, B, A+B) are materialised per plane-wave chunk by k_anl_gen3 into one real
// matrix W3: rows grouped by blocks of 8 projectors, row(p, kind) = (p/8)*24 + kind*8 + p%8, one double per plane wave,
// so that 8 consecutive rows are one m8 (k_fnl3) or two k4 (k_back3) DMMA operand slices of ONE kind.  fs is written in
// the same row order (F, H, F+H) by k_fnl_finish3.  The c operand stays interleaved (re,im): one LDS.128 yields X and Y
// of a fragment element, Y-X is one DADD per fragment element (2 % of the DMMA pipe time).
// Error: the Karatsuba imaginary part is accurate to eps * sum |anl||c| (normwise like the 4-product form; not
// componentwise) -- far inside the 1e-10 parity tolerance; summation stays deterministic (fixed split-K order).
#pragma once

namespace qb200 {

#define N3_KS 16                      // complex plane waves per stage of k_fnl3
#define N3_APITCH 20                  // doubles per A row in k_fnl3 (16 + 4: conflict-free LDS.64 fragments)
#define N3_BPITCH 40                  // doubles per B row in k_fnl3 (32 + 8: conflict-free LDS.128 fragments)
#define N3_BK_KROWS 24                // W3 rows per stage of k_back3: one block of 8 projectors x 3 kinds
#define N3_BK_BPITCH 36               // doubles per state row of the B tile of k_back3 (24 + 12, = 4 mod 16)

// Tile geometry: NWM x NWN warps; a warp owns 16 projectors (k_fnl3) or 16 plane waves (k_back3) x 32 states x 3 kinds
// (24 m8n8 accumulators).  <4,4>: 512 threads, one CTA per SM, 3-4 stages; <4,2>: 256 threads, two CTAs per SM whose
// stage barriers and pipeline fills overlap each other.
template <int NWM, int NWN, int NSTG> struct Fnl3Cfg {
  static constexpr int NTHR = 32 * NWM * NWN, MP = 16 * NWM, NT = 32 * NWN;
  static constexpr int ASTAGE = 3 * MP * N3_APITCH, STAGE = ASTAGE + NT * N3_BPITCH, SMEM = NSTG * STAGE * 8;
};
template <int NWM, int NWN, int NSTG> struct Back3Cfg {
  static constexpr int NTHR = 32 * NWM * NWN, GT = 16 * NWM, NT = 32 * NWN;
  static constexpr int APITCH = GT + 4, ASTAGE = N3_BK_KROWS * APITCH, STAGE = ASTAGE + NT * N3_BK_BPITCH, SMEM = NSTG * STAGE * 8;
};

// grid (ceil(gpad/128), na), block 128: as k_anl_gen, three rows (A, B, A+B) of one double per plane wave
__global__ void __launch_bounds__(128) k_anl_gen3(NlSpecies S, NlLattice L, int ngw, const double* __restrict__ kpgx, int gbeg,
                                                  int gcount, int gpad, double* __restrict__ W3, size_t WP)
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
    const int p = S.poff + ia * S.npr + ipr;
    double2 a = make_double2(0.0, 0.0);
    if (ok) a = anl_value(S.lproj[ipr], S.twnl[(size_t)ipr * ngw + g], sn, cs);
    double* row = W3 + ((size_t)(p >> 3) * 24 + (p & 7)) * WP + gl;
    row[0] = a.x;
    row[8 * WP] = a.y;
    row[16 * WP] = a.x + a.y;
  }
}

// ------------------------------------------------------------------------------------------------ fnl = anl^H c  (3M)
// grid (ceil(Mtot/MP), ceil(nst/NT), ksplit).  Reduction over the chunk's plane waves [blockIdx.z*kper, +kper) (kper a
// multiple of 16).  W3 needs no bounds checks: its pad rows (projectors >= Mtot of the last block of 8) and pad columns
// (plane waves gcount..gpad) are zero.  part[(ks*2*nst + 2n+{re,im})*Mp + p]  (=, or += when accumulate).
template <int NWM, int NWN, int NSTG>
__global__ void __launch_bounds__(32 * NWM * NWN, 512 / (32 * NWM * NWN))
k_fnl3(const double* __restrict__ W3, size_t WP, int gbeg, int gcount, int kper, const double2* __restrict__ c, size_t ldc, int nst,
       double* __restrict__ part, int Mp, int Mtot, int accumulate, int lower_only = 0)
{
  typedef Fnl3Cfg<NWM, NWN, NSTG> C;
  // lower_only (qb200_gram: the Hermitian overlap matrix): tiles that lie strictly above the diagonal are never read
  if (lower_only && (int)blockIdx.x * C::MP + C::MP - 1 < (int)blockIdx.y * C::NT) return;
  extern __shared__ __align__(16) double nl_smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp / NWN, wn = warp - wm * NWN;
  const int p0 = blockIdx.x * C::MP, n0 = blockIdx.y * C::NT;
  const int kbeg = blockIdx.z * kper, kend = min(kbeg + kper, gcount);
  const int nstage = kend > kbeg ? (kend - kbeg + N3_KS - 1) / N3_KS : 0;
  // A copies: 3*MP rows x 8 chunks (2 plane waves each); B copies: NT states x 16 plane waves
  constexpr int ACOP = 3 * C::MP * 8 / C::NTHR, BCOP = C::NT * 16 / C::NTHR, AROWS = C::NTHR / 8, BROWS = C::NTHR / 16;
  const int arow = tid >> 3, ach = tid & 7, bn = tid >> 4, bch = tid & 15;
  const double* wsrc = W3 + ((size_t)(p0 / 8) * 24 + arow) * WP + 2 * ach;
  const double2* csrc = c + (size_t)(n0 + bn) * ldc + gbeg + bch;
  auto issue = [&](int st) {
    if (st < nstage) {
      double* As = nl_smem + (st % NSTG) * C::STAGE;
      double* Bs = As + C::ASTAGE;
      const int k0 = kbeg + st * N3_KS;
#pragma unroll
      for (int i = 0; i < ACOP; i++)
        nl_cp16(As + (arow + AROWS * i) * N3_APITCH + 2 * ach, wsrc + (size_t)(AROWS * i) * WP + k0, true);
      const bool kok = k0 + bch < kend;
#pragma unroll
      for (int i = 0; i < BCOP; i++) {
        const bool ok = kok && n0 + bn + BROWS * i < nst;
        nl_cp16(Bs + (bn + BROWS * i) * N3_BPITCH + 2 * bch, ok ? csrc + (size_t)(BROWS * i) * ldc + k0 : c, ok);
      }
    }
    nl_cp_commit();
  };
  double acc[3][2][4][2];
#pragma unroll
  for (int q = 0; q < 3; q++)
#pragma unroll
    for (int i = 0; i < 2; i++)
#pragma unroll
      for (int j = 0; j < 4; j++) acc[q][i][j][0] = acc[q][i][j][1] = 0.0;
  const int r = lane >> 2, kq = lane & 3;
  for (int s = 0; s < NSTG - 1; s++) issue(s);
  for (int st = 0; st < nstage; st++) {
    nl_cp_wait_group<NSTG - 2>();
    __syncthreads();
    const double* As = nl_smem + (st % NSTG) * C::STAGE;
    const double* a0 = As + ((wm * 2) * 24 + r) * N3_APITCH + kq;
    const double* b0 = As + C::ASTAGE + (wn * 32 + r) * N3_BPITCH + 2 * kq;
    const bool work = !(p0 + wm * 16 >= Mtot || n0 + wn * 32 >= nst);   // false: warp tile entirely in the padding (warp-uniform)
    // the first step of the stage runs BEFORE the copies of the stage after next are issued: the tensor pipe gets work right
    // after the barrier instead of idling while every warp computes copy addresses (cf. warp_mma_stage in nonlocal.cu)
#pragma unroll
    for (int k4 = 0; k4 < N3_KS / 4; k4++) {
      if (k4 == 1) issue(st + NSTG - 1);
      if (!work) continue;
      double a[3][2];
#pragma unroll
      for (int q = 0; q < 3; q++)
#pragma unroll
        for (int i = 0; i < 2; i++) a[q][i] = a0[(i * 24 + q * 8) * N3_APITCH + k4 * 4];
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const double2 v = *reinterpret_cast<const double2*>(b0 + j * 8 * N3_BPITCH + k4 * 8);
        const double z = v.y - v.x;
#pragma unroll
        for (int i = 0; i < 2; i++) {
          dmma(acc[0][i][j][0], acc[0][i][j][1], a[0][i], v.x);
          dmma(acc[1][i][j][0], acc[1][i][j][1], a[1][i], v.y);
          dmma(acc[2][i][j][0], acc[2][i][j][1], a[2][i], z);
        }
      }
    }
  }
  const int ncols = 2 * nst;
#pragma unroll
  for (int i = 0; i < 2; i++)
#pragma unroll
    for (int j = 0; j < 4; j++)
#pragma unroll
      for (int e = 0; e < 2; e++) {
        const int p = p0 + wm * 16 + i * 8 + r;
        const int n = n0 + wn * 32 + j * 8 + 2 * kq + e;
        if (p < Mtot && n < nst) {
          const double P1 = acc[0][i][j][e], P2 = acc[1][i][j][e], P3 = acc[2][i][j][e];
          const double re = P1 + P2, im = P3 + (P1 - P2);
          double* dr = part + ((size_t)blockIdx.z * ncols + 2 * n) * Mp + p;
          double* di = dr + Mp;
          *dr = accumulate ? *dr + re : re;
          *di = accumulate ? *di + im : im;
        }
      }
}

// one thread per (n, p): split-K reduce, E_nl block partials, fs3[n][row(p,kind)] = wt/omega * (F, H, F+H)
__global__ void __launch_bounds__(256) k_fnl_finish3(const double* __restrict__ wtp, int Mtot, const double* __restrict__ part, int Mp,
                                                     int nst, int ksplit, const double* __restrict__ occ, double omega_inv,
                                                     double* __restrict__ fs3, int FP3, double* __restrict__ eblk)
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
    double* o = fs3 + (size_t)n * FP3 + (p >> 3) * 24 + (p & 7);
    const double F = fac * fr, Hh = fac * fi;
    o[0] = F; o[8] = Hh; o[16] = F + Hh;
  }
  red[threadIdx.x] = e;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) eblk[blockIdx.x] = red[0];
}

// ------------------------------------------------------------------------------------------------ cp += anl * fs  (3M)
// grid (ceil(nst/NT), ceil(gcount/GT)) -- state tiles fastest, so the CTAs that share a W3 tile run together and W3
// streams from HBM once.  Reduction over the RW3 = 24*ceil(Mtot/8) rows of W3, 24 (one projector block, all kinds) per
// stage; pad rows/columns of W3 and pad rows of fs3 are zero.
template <int NWM, int NWN, int NSTG>
__global__ void __launch_bounds__(32 * NWM * NWN, 512 / (32 * NWM * NWN))
k_back3(const double* __restrict__ W3, size_t WP, int RW3, int gbeg, int gcount, const double* __restrict__ fs3, int FP3,
        double2* __restrict__ cp, size_t ldc, int nst, int overwrite, int upper_tri = 0)
{
  typedef Back3Cfg<NWM, NWN, NSTG> C;
  extern __shared__ __align__(16) double nl_smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp / NWN, wn = warp - wm * NWN;
  const int n0 = blockIdx.x * C::NT, gl0 = blockIdx.y * C::GT;
  // upper_tri (qb200_gram: the second operand L^-H is upper triangular): rows m > n0 + NT - 1 contribute nothing
  const int nstage = upper_tri ? min(RW3 / N3_BK_KROWS, (n0 + C::NT + 7) / 8) : RW3 / N3_BK_KROWS;
  constexpr int ACH = C::GT / 2, ATOT = N3_BK_KROWS * ACH, BTOT = C::NT * 12;   // 16-byte copies per stage
  auto issue = [&](int st) {
    if (st < nstage) {
      double* As = nl_smem + (st % NSTG) * C::STAGE;
      double* Bs = As + C::ASTAGE;
      const int k0 = st * N3_BK_KROWS;
#pragma unroll
      for (int i = 0; i < (ATOT + C::NTHR - 1) / C::NTHR; i++) {
        const int ci = tid + i * C::NTHR;
        if (ATOT % C::NTHR == 0 || ci < ATOT) {
          const int kr = ci / ACH, gc = ci - kr * ACH;
          nl_cp16(As + kr * C::APITCH + 2 * gc, W3 + (size_t)(k0 + kr) * WP + gl0 + 2 * gc, true);
        }
      }
#pragma unroll
      for (int i = 0; i < BTOT / C::NTHR; i++) {
        const int ci = tid + i * C::NTHR, nl = ci / 12, ch = ci - nl * 12;
        const bool ok = n0 + nl < nst;
        nl_cp16(Bs + nl * N3_BK_BPITCH + 2 * ch, ok ? fs3 + (size_t)(n0 + nl) * FP3 + k0 + 2 * ch : fs3, ok);
      }
    }
    nl_cp_commit();
  };
  double acc[3][2][4][2];
#pragma unroll
  for (int q = 0; q < 3; q++)
#pragma unroll
    for (int i = 0; i < 2; i++)
#pragma unroll
      for (int j = 0; j < 4; j++) acc[q][i][j][0] = acc[q][i][j][1] = 0.0;
  const int r = lane >> 2, kq = lane & 3;
  for (int s = 0; s < NSTG - 1; s++) issue(s);
  for (int st = 0; st < nstage; st++) {
    nl_cp_wait_group<NSTG - 2>();
    __syncthreads();
    const double* As = nl_smem + (st % NSTG) * C::STAGE;
    const double* a0 = As + kq * C::APITCH + wm * 16 + r;
    const double* b0 = As + C::ASTAGE + (wn * 32 + r) * N3_BK_BPITCH + kq;
#pragma unroll
    for (int q = 0; q < 3; q++)
#pragma unroll
      for (int k4 = 0; k4 < 2; k4++) {
        if (q == 0 && k4 == 1) issue(st + NSTG - 1);      // after the first step of the stage (see k_fnl3)
        double a[2], b[4];
#pragma unroll
        for (int i = 0; i < 2; i++) a[i] = a0[(q * 8 + k4 * 4) * C::APITCH + i * 8];
#pragma unroll
        for (int j = 0; j < 4; j++) b[j] = b0[j * 8 * N3_BK_BPITCH + q * 8 + k4 * 4];
#pragma unroll
        for (int i = 0; i < 2; i++)
#pragma unroll
          for (int j = 0; j < 4; j++) dmma(acc[q][i][j][0], acc[q][i][j][1], a[i], b[j]);
      }
  }
#pragma unroll
  for (int i = 0; i < 2; i++)
#pragma unroll
    for (int j = 0; j < 4; j++)
#pragma unroll
      for (int e = 0; e < 2; e++) {
        const int gl = gl0 + wm * 16 + i * 8 + r;
        const int n = n0 + wn * 32 + j * 8 + 2 * kq + e;
        if (gl < gcount && n < nst) {
          const double R1 = acc[0][i][j][e], R2 = acc[1][i][j][e], R3 = acc[2][i][j][e];
          double2* dst = cp + (size_t)n * ldc + gbeg + gl;
          double2 v = overwrite ? make_double2(0.0, 0.0) : *dst;     // overwrite: cp is known to be zero (H psi: first term)
          v.x += R1 - R2;
          v.y += R3 - (R1 + R2);
          *dst = v;
        }
      }
}

}  // namespace qb200
