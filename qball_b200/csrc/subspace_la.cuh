// qball_b200/csrc/subspace_la.cuh -- SURVEY section 8 row f1: the subspace dense linear algebra that sits between two
// H psi evaluations of a ground-state iteration, on the device and on the same FP64 tensor-core GEMM kernels as the
// projector contractions (this file is included at the end of nonlocal.cu and reuses k_fnl3 / k_back3 / k_fnl / k_back):
//
//   qb200_residual : A = C^H (HC) ; HC -= C A          PSDAWavefunctionStepper.cc:65-84 (real), :264-277 (complex);
//                                                       PSDWavefunctionStepper.cc:62-90 is the same sequence
//   qb200_gram     : S = C^H C ; S = L L^H ; C <- C L^-H   SlaterDet::gram, SlaterDet.cc:1043-1143 (norm-conserving)
//
// Both are "projector contractions with the wavefunctions as projectors": the block C plays the role of anl.
//   complex bases: C is packed per chunk of plane waves into the three-kind real matrix W3 of nonlocal_3m.cuh
//                  (rows grouped by 8 states: Re, Im, Re+Im), k_fnl3 forms C^H X as three real DMMA GEMMs, a finish
//                  kernel reduces the split-K partials in fixed order and writes the second operand in k_back3's row order
//   real bases   : the DoubleMatrix proxy of the reference (2*mloc real rows) is the block itself, so k_fnl<1> / k_back<0>
//                  read it in place as their W operand (residual) or a per-chunk copy (gram: C is overwritten);
//                  the factor 2 (G and -G) and the rank-1 term of real row 0 (ger / syr) are applied in the finish kernel.
// gram: the nst x nst Cholesky factorisation runs on the device as a blocked right-looking algorithm (32 x 32 diagonal
// blocks factorised and inverted in shared memory, panel and trailing updates as tile kernels), then L^-1 is formed by
// block forward substitution and C <- C L^-H is one more k_back3 / k_back GEMM (overwrite form) from the packed copy.
// Rows ig >= ngw of the blocks are padding (zero in the reference, SlaterDet.cc:2784-2787) and are neither read nor written.
#pragma once

namespace qb200 {

#define LA_NB 32

// ------------------------------------------------------------------------------------------------ packing
// grid (ceil(gpad/128), nall), block 128: state m, plane waves [gbeg, gbeg+gcount) of the chunk -> W3 rows (Re, Im, Re+Im)
__global__ void __launch_bounds__(128) k_pack_w3(const double2* __restrict__ c, size_t ldc, int gbeg, int gcount, int gpad,
                                                 double* __restrict__ W3, size_t WP)
{
  const int gl = blockIdx.x * 128 + threadIdx.x;
  if (gl >= gpad) return;
  const int m = blockIdx.y;
  double2 a = make_double2(0.0, 0.0);
  if (gl < gcount) a = c[(size_t)m * ldc + gbeg + gl];
  double* row = W3 + ((size_t)(m >> 3) * 24 + (m & 7)) * WP + gl;
  row[0] = a.x;
  row[8 * WP] = a.y;
  row[16 * WP] = a.x + a.y;
}
// real bases: W[m][2g..2g+1] = c[g, m] for the chunk (a copy, because gram overwrites c)
__global__ void __launch_bounds__(128) k_pack_wr(const double2* __restrict__ c, size_t ldc, int gbeg, int gcount, int gpad,
                                                 double* __restrict__ W, size_t WP)
{
  const int gl = blockIdx.x * 128 + threadIdx.x;
  if (gl >= gpad) return;
  const int m = blockIdx.y;
  double2 a = make_double2(0.0, 0.0);
  if (gl < gcount) a = c[(size_t)m * ldc + gbeg + gl];
  *reinterpret_cast<double2*>(W + (size_t)m * WP + 2 * gl) = a;
}

// ------------------------------------------------------------------------------------------------ finish: partials -> A or S
// one thread per (n, m), m < nall rows, n < nst columns.  part as written by k_fnl3 / k_fnl<1>.
//   RESIDUAL: a_out[n*nall + m] = A[m,n] (optional), second operand of the back GEMM = -A
//   GRAM    : S[n*nS + m] = S[m,n] as complex (both triangles; potrf reads the lower one)
// real bases: value = 2*sum - c(row 0, m) * x(row 0, n)   (gemm alpha 2.0 + ger/syr -1.0 on real row 0)
template <int IS_REAL, int GRAM>
__global__ void __launch_bounds__(256) k_la_finish(const double* __restrict__ part, int Mp, int nall, int nst, int ksplit,
                                                   const double2* __restrict__ c, const double2* __restrict__ x, size_t ldc,
                                                   double* __restrict__ a_out, double* __restrict__ fs, int FP,
                                                   double2* __restrict__ S, int nS)
{
  const size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (idx >= (size_t)nst * nall) return;
  const int n = (int)(idx / nall), m = (int)(idx % nall);
  const int ncols = IS_REAL ? nst : 2 * nst;
  double fr = 0.0, fi = 0.0;
  if (IS_REAL) {
    for (int ks = 0; ks < ksplit; ks++) fr += part[((size_t)ks * ncols + n) * Mp + m];
    fr = 2.0 * fr - c[(size_t)m * ldc].x * x[(size_t)n * ldc].x;
  } else {
    for (int ks = 0; ks < ksplit; ks++) {
      fr += part[((size_t)ks * ncols + 2 * n) * Mp + m];
      fi += part[((size_t)ks * ncols + 2 * n + 1) * Mp + m];
    }
  }
  if (GRAM) {
    S[(size_t)n * nS + m] = make_double2(fr, fi);
  } else if (IS_REAL) {
    if (a_out) a_out[(size_t)n * nall + m] = fr;
    fs[(size_t)n * FP + m] = -fr;
  } else {
    if (a_out) { a_out[2 * ((size_t)n * nall + m)] = fr; a_out[2 * ((size_t)n * nall + m) + 1] = fi; }
    double* o = fs + (size_t)n * FP + (m >> 3) * 24 + (m & 7);
    o[0] = -fr; o[8] = -fi; o[16] = -(fr + fi);
  }
}

// ------------------------------------------------------------------------------------------------ blocked Cholesky
__device__ __forceinline__ double2 cmulc(double2 a, double2 b) { return make_double2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y); }  // a conj(b)

// one CTA (32 x 32): factorise the diagonal block at k0 in shared memory (unblocked, LAPACK zpotf2's recurrence),
// write L_kk back and its inverse (lower triangular) to Dinv (32 x 32 column-major, zero upper part).
// info (device int): first non-positive pivot + 1, as LAPACK.
__global__ void __launch_bounds__(1024) k_potrf_diag(double2* __restrict__ S, int n, int k0, double2* __restrict__ Dinv,
                                                     int* __restrict__ info)
{
  __shared__ double2 A[LA_NB][LA_NB + 1];     // A[col][row]
  __shared__ double2 X[LA_NB][LA_NB + 1];
  const int i = threadIdx.x, j = threadIdx.y;
  const int nb = min(LA_NB, n - k0);
  double2 v = make_double2(i == j ? 1.0 : 0.0, 0.0);          // identity padding beyond the matrix
  if (i < nb && j < nb) v = i >= j ? S[(size_t)(k0 + j) * n + k0 + i] : make_double2(0.0, 0.0);
  A[j][i] = v;
  X[j][i] = make_double2(0.0, 0.0);
  __syncthreads();
  for (int p = 0; p < LA_NB; p++) {
    if (i == p && j == p) {
      double d = A[p][p].x;
      if (!(d > 0.0)) { atomicCAS(info, 0, k0 + p + 1); d = 1.0; }
      A[p][p] = make_double2(sqrt(d), 0.0);
    }
    __syncthreads();
    if (j == p && i > p) { const double dinv = 1.0 / A[p][p].x; A[p][i] = make_double2(A[p][i].x * dinv, A[p][i].y * dinv); }
    __syncthreads();
    if (j > p && i >= j) { const double2 t = cmulc(A[p][i], A[p][j]); A[j][i].x -= t.x; A[j][i].y -= t.y; }
    __syncthreads();
  }
  if (i < nb && j < nb && i >= j) S[(size_t)(k0 + j) * n + k0 + i] = A[j][i];
  // inverse of the lower triangular block by forward substitution, one thread per column
  if (j == 0) {
    const int col = i;
    for (int r = col; r < LA_NB; r++) {
      double2 s = make_double2(r == col ? 1.0 : 0.0, 0.0);
      for (int k = col; k < r; k++) { const double2 t = cmul(A[k][r], X[col][k]); s.x -= t.x; s.y -= t.y; }
      const double dinv = 1.0 / A[r][r].x;
      X[col][r] = make_double2(s.x * dinv, s.y * dinv);
    }
  }
  __syncthreads();
  Dinv[j * LA_NB + i] = i >= j ? X[j][i] : make_double2(0.0, 0.0);
}

// grid (row blocks below the diagonal block): A[r0.., k0..] <- A[r0.., k0..] * L_kk^-H = A * Dinv^H
__global__ void __launch_bounds__(1024) k_potrf_panel(double2* __restrict__ S, int n, int k0, const double2* __restrict__ Dinv)
{
  __shared__ double2 A[LA_NB][LA_NB + 1];     // A[col][row]
  __shared__ double2 D[LA_NB][LA_NB + 1];     // D[col][row] = Dinv
  const int i = threadIdx.x, j = threadIdx.y;
  const int r = k0 + LA_NB + blockIdx.x * LA_NB + i;
  A[j][i] = (r < n && k0 + j < n) ? S[(size_t)(k0 + j) * n + r] : make_double2(0.0, 0.0);
  D[j][i] = Dinv[j * LA_NB + i];
  __syncthreads();
  double2 s = make_double2(0.0, 0.0);
  for (int p = 0; p <= j; p++) { const double2 t = cmulc(A[p][i], D[p][j]); s.x += t.x; s.y += t.y; }   // sum_p A[i,p] conj(Dinv[j,p])
  if (r < n && k0 + j < n) S[(size_t)(k0 + j) * n + r] = s;
}

// grid (nt, nt), nt = trailing row blocks: S[bi, bj] -= P_bi P_bj^H for bi >= bj (P = the panel just computed)
__global__ void __launch_bounds__(1024) k_potrf_trail(double2* __restrict__ S, int n, int k0)
{
  if (blockIdx.x < blockIdx.y) return;
  __shared__ double2 Pi[LA_NB][LA_NB + 1];    // [p][row]
  __shared__ double2 Pj[LA_NB][LA_NB + 1];
  const int i = threadIdx.x, j = threadIdx.y;
  const int t0 = k0 + LA_NB;
  const int ri = t0 + blockIdx.x * LA_NB + i, rj = t0 + blockIdx.y * LA_NB + i;
  const bool pok = k0 + j < n;
  Pi[j][i] = (ri < n && pok) ? S[(size_t)(k0 + j) * n + ri] : make_double2(0.0, 0.0);
  Pj[j][i] = (rj < n && pok) ? S[(size_t)(k0 + j) * n + rj] : make_double2(0.0, 0.0);
  __syncthreads();
  double2 s = make_double2(0.0, 0.0);
#pragma unroll 8
  for (int p = 0; p < LA_NB; p++) { const double2 t = cmulc(Pi[p][i], Pj[p][j]); s.x += t.x; s.y += t.y; }
  const int cj = t0 + blockIdx.y * LA_NB + j;
  if (ri < n && cj < n && ri >= cj) {
    double2* d = S + (size_t)cj * n + ri;
    d->x -= s.x; d->y -= s.y;
  }
}

// grid (nblk): block column k of X = L^-1 by block forward substitution:
//   X[k,k] = Dinv_k ; X[i,k] = -Dinv_i * sum_{j=k}^{i-1} L[i,j] X[j,k]      (X zero above the diagonal, pre-cleared)
__global__ void __launch_bounds__(1024) k_trtri_cols(const double2* __restrict__ L, int n, const double2* __restrict__ Dinv,
                                                     double2* __restrict__ X)
{
  __shared__ double2 Acc[LA_NB][LA_NB + 1];   // [col][row]
  __shared__ double2 D[LA_NB][LA_NB + 1];
  const int r = threadIdx.x, cc = threadIdx.y;
  const int k = blockIdx.x, nblk = (n + LA_NB - 1) / LA_NB;
  const int col = k * LA_NB + cc;
  if (col < n && k * LA_NB + r < n) X[(size_t)col * n + k * LA_NB + r] = Dinv[(size_t)k * LA_NB * LA_NB + cc * LA_NB + r];
  for (int i = k + 1; i < nblk; i++) {
    __syncthreads();                            // X[j,k] of the previous steps visible; Acc / D free
    const int row = i * LA_NB + r;
    double2 s = make_double2(0.0, 0.0);
    if (row < n && col < n)
      for (int q = k * LA_NB; q < i * LA_NB; q++) {
        const double2 t = cmul(L[(size_t)q * n + row], X[(size_t)col * n + q]);
        s.x += t.x; s.y += t.y;
      }
    Acc[cc][r] = s;
    D[cc][r] = Dinv[(size_t)i * LA_NB * LA_NB + cc * LA_NB + r];
    __syncthreads();
    double2 o = make_double2(0.0, 0.0);
    for (int p = 0; p <= r; p++) { const double2 t = cmul(D[p][r], Acc[cc][p]); o.x -= t.x; o.y -= t.y; }
    if (row < n && col < n) X[(size_t)col * n + row] = o;
  }
}

// second operand of the back GEMM of gram: T = L^-H, T[m,n] = conj(X[n,m]) (upper triangular)
template <int IS_REAL>
__global__ void __launch_bounds__(256) k_gram_operand(const double2* __restrict__ X, int n, double* __restrict__ fs, int FP)
{
  const size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (idx >= (size_t)n * n) return;
  const int col = (int)(idx / n), m = (int)(idx % n);
  double2 t = make_double2(0.0, 0.0);
  if (m <= col) { const double2 x = X[(size_t)m * n + col]; t = make_double2(x.x, -x.y); }
  if (IS_REAL) fs[(size_t)col * FP + m] = t.x;
  else {
    double* o = fs + (size_t)col * FP + (m >> 3) * 24 + (m & 7);
    o[0] = t.x; o[8] = t.y; o[16] = t.x + t.y;
  }
}

// the same for the columns [first, first + ncol) of T only (band-sharded gram: this rank's states)
template <int IS_REAL>
__global__ void __launch_bounds__(256) k_gram_operand_cols(const double2* __restrict__ X, int n, int first, int ncol, double* __restrict__ fs, int FP)
{
  const size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (idx >= (size_t)ncol * n) return;
  const int cl = (int)(idx / n), m = (int)(idx % n), col = first + cl;
  double2 t = make_double2(0.0, 0.0);
  if (m <= col) { const double2 x = X[(size_t)m * n + col]; t = make_double2(x.x, -x.y); }
  if (IS_REAL) fs[(size_t)cl * FP + m] = t.x;
  else {
    double* o = fs + (size_t)cl * FP + (m >> 3) * 24 + (m & 7);
    o[0] = t.x; o[8] = t.y; o[16] = t.x + t.y;
  }
}

}  // namespace qb200

using namespace qb200;

struct qb200_la {
  int device;
  cudaStream_t stream;
  int ngw, is_real;
  long long budget;                             // bytes the packed copy W may take
  double *W, *part, *fs; size_t W_cap, part_cap, fs_cap;
  size_t W_WP;                                  // pitch W was last zero-filled for (3M pad rows must be zero)
  double *S, *X, *Dinv; size_t S_cap, X_cap, Dinv_cap;
  bool jacobi_graph;                            // the last qb200_diag replayed its sweeps from a CUDA graph
  bool jacobi_blocked;                          // the last qb200_diag ran the blocked Jacobi method
  double* Ssh; size_t Ssh_cap;                  // band-sharded gram: the overlap columns before / after the sum over ranks
  double *st_c, *st_x, *st_a; size_t st_c_cap, st_x_cap, st_a_cap;
  int* info_dev;
  long long launches;
  int nsm;
  int nchunks_last;
};

extern "C" int qb200_la_create(qb200_la** out, int device, int ngw, int is_real)
{
  if (!out || ngw < 1) { set_error("qb200_la_create: bad argument"); return QB200_EINVAL; }
  *out = nullptr;
  int ndev = 0;
  QB_CUDA(cudaGetDeviceCount(&ndev));
  if (device < 0 || device >= ndev) { set_error("qb200_la_create: no such CUDA device"); return QB200_ENODEV; }
  QB_CUDA(cudaSetDevice(device));
  qb200_la* la = new qb200_la();
  la->device = device; la->stream = 0; la->ngw = ngw; la->is_real = is_real ? 1 : 0;
  la->budget = 8ll << 30;
  if (const char* e = getenv("QB200_LA_BYTES")) la->budget = std::max(1ll << 20, atoll(e));
  la->W = la->part = la->fs = la->S = la->X = la->Dinv = la->st_c = la->st_x = la->st_a = nullptr;
  la->Ssh = nullptr; la->Ssh_cap = 0; la->jacobi_graph = false; la->jacobi_blocked = false;
  la->W_cap = la->part_cap = la->fs_cap = la->S_cap = la->X_cap = la->Dinv_cap = la->st_c_cap = la->st_x_cap = la->st_a_cap = 0;
  la->W_WP = 0; la->launches = 0; la->nchunks_last = 0; la->info_dev = nullptr;
  cudaDeviceProp prop;
  QB_CUDA(cudaGetDeviceProperties(&prop, device));
  la->nsm = prop.multiProcessorCount;
  QB_CUDA(cudaMalloc((void**)&la->info_dev, sizeof(int)));
  QB_CUDA(cudaFuncSetAttribute(k_fnl<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, FNL_SMEM_BYTES));
  QB_CUDA(cudaFuncSetAttribute(k_back<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, BK_SMEM_BYTES));
  QB_CUDA(cudaFuncSetAttribute(k_fnl3<4, 2, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, Fnl3Cfg<4, 2, 2>::SMEM));
  QB_CUDA(cudaFuncSetAttribute(k_back3<4, 2, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, Back3Cfg<4, 2, 3>::SMEM));
  *out = la;
  return QB200_OK;
}

extern "C" int qb200_la_set_stream(qb200_la* la, void* s)
{
  if (!la) return QB200_EINVAL;
  la->stream = (cudaStream_t)s;
  return QB200_OK;
}

extern "C" int qb200_la_set_workspace(qb200_la* la, long long bytes)
{
  if (!la || bytes < (1ll << 20)) { set_error("qb200_la_set_workspace: bad argument"); return QB200_EINVAL; }
  la->budget = bytes;
  return QB200_OK;
}

extern "C" int qb200_la_destroy(qb200_la* la)
{
  if (!la) return QB200_OK;
  cudaSetDevice(la->device);
  for (double* p : { la->W, la->part, la->fs, la->S, la->X, la->Dinv, la->Ssh, la->st_c, la->st_x, la->st_a }) if (p) cudaFree(p);
  if (la->info_dev) cudaFree(la->info_dev);
  delete la;
  return QB200_OK;
}

extern "C" long long qb200_la_query(const qb200_la* la, int what)
{
  if (!la) return -1;
  switch (what) {
    case 9: return la->launches;
    case 11: return la->nchunks_last;
    case 12: return (long long)(la->W_cap * sizeof(double));
    case 13: return la->jacobi_graph ? 1 : 0;
    case 14: return la->jacobi_blocked ? 1 : 0;
    default: return -1;
  }
}

namespace {

struct LaGeom {
  bool m3;                 // complex: three-kind packed copy + k_fnl3 / k_back3
  int nall, RW, Mp, FP;    // rows of W, even pitch of part, pitch of the second operand
  int gchunk, nchunks;
  size_t WP, Welems;
  int mt, nt, ksplit;
};

// nall "projector" rows, nst columns; in_place: real bases read the block itself as W (no copy, one chunk)
// lower_only: the first GEMM computes a Hermitian matrix and its CTAs above the diagonal exit at once (k_fnl3) -- the split-K
// that fills the SMs is chosen for the tiles that actually work
LaGeom la_geometry(const qb200_la* la, int nall, int nst, bool in_place, size_t ldc, bool lower_only = false)
{
  LaGeom g;
  g.m3 = !la->is_real;
  g.nall = nall;
  g.RW = g.m3 ? 24 * ((nall + 7) / 8) : nall;
  g.Mp = (nall + 1) & ~1;
  g.FP = g.m3 ? g.RW : g.Mp;
  const int ngw = la->ngw;
  if (in_place) {
    g.gchunk = (ngw + 15) / 16 * 16; g.nchunks = 1; g.WP = 2 * ldc; g.Welems = 0;
  } else {
    const long long per_g = g.m3 ? 8ll * 192 * ((nall + 63) / 64) : 16ll * nall;
    long long gmax = la->budget / std::max(per_g, 1ll);
    gmax = std::max(512ll, (gmax / 512) * 512);
    g.gchunk = (int)std::min<long long>(gmax, ((long long)ngw + 15) / 16 * 16);
    g.nchunks = (ngw + g.gchunk - 1) / g.gchunk;
    g.WP = g.m3 ? (size_t)g.gchunk : 2 * (size_t)g.gchunk;
    g.Welems = g.m3 ? (size_t)192 * ((nall + 63) / 64) * g.WP + 128 : (size_t)g.RW * g.WP;
  }
  g.mt = g.m3 ? (nall + 63) / 64 : (g.RW + NL_TM - 1) / NL_TM;
  g.nt = g.m3 ? (nst + 63) / 64 : (nst + NL_TN - 1) / NL_TN;
  const int slots = g.m3 ? 2 * la->nsm : la->nsm;
  g.ksplit = 1;
  const int maxk = std::max(1, std::min(g.gchunk, ngw) / 512);
  double best = -1.0;
  long tiles = (long)g.mt * g.nt;
  if (lower_only && g.m3) {
    tiles = 0;
    for (int bx = 0; bx < g.mt; bx++)
      for (int by = 0; by < g.nt; by++) if (!(bx * 64 + 63 < by * 64)) tiles++;
  }
  for (int k = 1; k <= std::min(maxk, 64); k++) {
    const long ctas = tiles * k;
    const long waves = (ctas + slots - 1) / slots;
    const double eff = (double)ctas / (double)(waves * slots) - 0.002 * k;
    if (eff > best) { best = eff; g.ksplit = k; }
  }
  return g;
}

#define LA_LAUNCH_CHECK(la) do { (la)->launches++; cudaError_t e__ = cudaGetLastError(); \
    if (e__ != cudaSuccess) return qb200::cuda_fail(e__, "kernel launch", __FILE__, __LINE__); } while (0)

int la_prepare_W(qb200_la* la, const LaGeom& g)
{
  int rc;
  if (g.Welems == 0) return QB200_OK;
  if (la->W_cap < g.Welems) la->W_WP = 0;
  if ((rc = nl_ensure(&la->W, &la->W_cap, g.Welems))) return rc;
  if (g.m3 && la->W_WP != g.WP) {                    // pad rows of the last block of 8 / tile of 64 must be zero
    QB_CUDA(cudaMemsetAsync(la->W, 0, la->W_cap * sizeof(double), la->stream));
    la->W_WP = g.WP;
  }
  return QB200_OK;
}

int la_pack(qb200_la* la, const LaGeom& g, const double* c, size_t ldc, int gbeg, int gcount, int gpad)
{
  dim3 grid((gpad + 127) / 128, g.nall);
  prof_begin(7, la->stream);
  if (g.m3) k_pack_w3<<<grid, 128, 0, la->stream>>>((const double2*)c, ldc, gbeg, gcount, gpad, la->W, g.WP);
  else k_pack_wr<<<grid, 128, 0, la->stream>>>((const double2*)c, ldc, gbeg, gcount, gpad, la->W, g.WP);
  prof_end(la->stream);
  LA_LAUNCH_CHECK(la);
  return QB200_OK;
}

// part (+)= W^H x over one chunk
int la_fnl(qb200_la* la, const LaGeom& g, const double* W, int gbeg, int gcount, const double* x, size_t ldc, int nst, bool accumulate,
           int lower_only = 0)
{
  int kper = ((g.m3 ? 1 : 2) * gcount + g.ksplit - 1) / g.ksplit;
  kper = g.m3 ? (kper + N3_KS - 1) / N3_KS * N3_KS : (kper + NL_KSTEP - 1) / NL_KSTEP * NL_KSTEP;
  dim3 g1(g.mt, g.nt, g.ksplit);
  prof_begin(3, la->stream);
  if (g.m3) k_fnl3<4, 2, 2><<<g1, 256, Fnl3Cfg<4, 2, 2>::SMEM, la->stream>>>(W, g.WP, gbeg, gcount, kper, (const double2*)x, ldc, nst, la->part, g.Mp, g.nall, accumulate, lower_only);
  else k_fnl<1><<<g1, NL_THREADS, FNL_SMEM_BYTES, la->stream>>>(W, g.WP, g.RW, gbeg, gcount, kper, (const double2*)x, ldc, nst, la->part, g.Mp, g.nall, accumulate);
  prof_end(la->stream);
  LA_LAUNCH_CHECK(la);
  return QB200_OK;
}

// y[rows of the chunk, :] (+)= W * fs
int la_back(qb200_la* la, const LaGeom& g, const double* W, int gbeg, int gcount, double* y, size_t ldc, int nst, int overwrite,
            int upper_tri = 0)
{
  prof_begin(5, la->stream);
  if (g.m3) k_back3<4, 2, 3><<<dim3(g.nt, (gcount + 63) / 64), 256, Back3Cfg<4, 2, 3>::SMEM, la->stream>>>(W, g.WP, g.RW, gbeg, gcount, la->fs, g.FP, (double2*)y, ldc, nst, overwrite, upper_tri);
  else k_back<0><<<dim3(g.nt, (gcount + 63) / 64), NL_THREADS, BK_SMEM_BYTES, la->stream>>>(W, g.WP, g.RW, gbeg, gcount, la->fs, g.FP, (double2*)y, ldc, nst, overwrite);
  prof_end(la->stream);
  LA_LAUNCH_CHECK(la);
  return QB200_OK;
}

}  // namespace

// device pointers
static int la_residual_dev(qb200_la* la, int ldc, int nall, const double* c, int nst, double* hc, double* a)
{
  int rc;
  const bool in_place = la->is_real != 0;            // real bases: the block is its own W operand
  const LaGeom g = la_geometry(la, nall, nst, in_place, ldc);
  la->nchunks_last = g.nchunks;
  if ((rc = la_prepare_W(la, g))) return rc;
  const int ncols = la->is_real ? nst : 2 * nst;
  if ((rc = nl_ensure(&la->part, &la->part_cap, (size_t)g.ksplit * ncols * g.Mp))) return rc;
  if ((rc = nl_ensure(&la->fs, &la->fs_cap, (size_t)nst * g.FP))) return rc;
  if (g.FP != (g.m3 ? 3 * nall : nall)) QB_CUDA(cudaMemsetAsync(la->fs, 0, (size_t)nst * g.FP * sizeof(double), la->stream));
  const int ngw = la->ngw;
  for (int ch = 0; ch < g.nchunks; ch++) {
    const int gbeg = ch * g.gchunk, gcount = std::min(g.gchunk, ngw - gbeg), gpad = (gcount + 15) / 16 * 16;
    if (!in_place && (rc = la_pack(la, g, c, ldc, gbeg, gcount, gpad))) return rc;
    if ((rc = la_fnl(la, g, in_place ? c : la->W, gbeg, gcount, hc, ldc, nst, ch > 0))) return rc;
  }
  const size_t total = (size_t)nst * nall;
  const int nblk = (int)((total + 255) / 256);
  prof_begin(4, la->stream);
  if (la->is_real) k_la_finish<1, 0><<<nblk, 256, 0, la->stream>>>(la->part, g.Mp, nall, nst, g.ksplit, (const double2*)c, (const double2*)hc, ldc, a, la->fs, g.FP, nullptr, 0);
  else k_la_finish<0, 0><<<nblk, 256, 0, la->stream>>>(la->part, g.Mp, nall, nst, g.ksplit, (const double2*)c, (const double2*)hc, ldc, a, la->fs, g.FP, nullptr, 0);
  prof_end(la->stream);
  LA_LAUNCH_CHECK(la);
  for (int i = 0; i < g.nchunks; i++) {
    const int ch = g.nchunks - 1 - i;
    const int gbeg = ch * g.gchunk, gcount = std::min(g.gchunk, ngw - gbeg), gpad = (gcount + 15) / 16 * 16;
    if (!in_place && i > 0 && (rc = la_pack(la, g, c, ldc, gbeg, gcount, gpad))) return rc;   // i == 0: still packed
    if ((rc = la_back(la, g, in_place ? c : la->W, gbeg, gcount, hc, ldc, nst, 0))) return rc;
  }
  return QB200_OK;
}

// la->S (n x n, lower triangle read) = L L^H, blocked right-looking, Dinv[k] = L_kk^-1; then la->X = L^-1
static int la_factor(qb200_la* la, int n)
{
  const int nblkd = (n + LA_NB - 1) / LA_NB;
  double2* S = (double2*)la->S; double2* Dinv = (double2*)la->Dinv;
  const dim3 tb(LA_NB, LA_NB);
  for (int k = 0; k < nblkd; k++) {
    const int k0 = k * LA_NB;
    k_potrf_diag<<<1, tb, 0, la->stream>>>(S, n, k0, Dinv + (size_t)k * LA_NB * LA_NB, la->info_dev);
    LA_LAUNCH_CHECK(la);
    const int nt = nblkd - k - 1;
    if (nt > 0) {
      k_potrf_panel<<<nt, tb, 0, la->stream>>>(S, n, k0, Dinv + (size_t)k * LA_NB * LA_NB);
      LA_LAUNCH_CHECK(la);
      k_potrf_trail<<<dim3(nt, nt), tb, 0, la->stream>>>(S, n, k0);
      LA_LAUNCH_CHECK(la);
    }
  }
  k_trtri_cols<<<nblkd, tb, 0, la->stream>>>(S, n, Dinv, (double2*)la->X);
  LA_LAUNCH_CHECK(la);
  return QB200_OK;
}

static int la_gram_dev(qb200_la* la, int ldc, int nst, double* c, int* info)
{
  int rc;
  const int n = nst;
  const LaGeom g = la_geometry(la, n, n, false, ldc, true);
  la->nchunks_last = g.nchunks;
  if ((rc = la_prepare_W(la, g))) return rc;
  const int ncols = la->is_real ? n : 2 * n;
  const int nblkd = (n + LA_NB - 1) / LA_NB;
  if ((rc = nl_ensure(&la->part, &la->part_cap, (size_t)g.ksplit * ncols * g.Mp))) return rc;
  if ((rc = nl_ensure(&la->fs, &la->fs_cap, (size_t)n * g.FP))) return rc;
  if ((rc = nl_ensure(&la->S, &la->S_cap, 2 * (size_t)n * n))) return rc;
  if ((rc = nl_ensure(&la->X, &la->X_cap, 2 * (size_t)n * n))) return rc;
  if ((rc = nl_ensure(&la->Dinv, &la->Dinv_cap, 2 * (size_t)nblkd * LA_NB * LA_NB))) return rc;
  QB_CUDA(cudaMemsetAsync(la->fs, 0, (size_t)n * g.FP * sizeof(double), la->stream));
  QB_CUDA(cudaMemsetAsync(la->X, 0, 2 * (size_t)n * n * sizeof(double), la->stream));
  QB_CUDA(cudaMemsetAsync(la->info_dev, 0, sizeof(int), la->stream));
  const int ngw = la->ngw;
  // S = C^H C
  for (int ch = 0; ch < g.nchunks; ch++) {
    const int gbeg = ch * g.gchunk, gcount = std::min(g.gchunk, ngw - gbeg), gpad = (gcount + 15) / 16 * 16;
    if ((rc = la_pack(la, g, c, ldc, gbeg, gcount, gpad))) return rc;
    if ((rc = la_fnl(la, g, la->W, gbeg, gcount, c, ldc, n, ch > 0, 1))) return rc;   // Hermitian: lower tiles only
  }
  const size_t total = (size_t)n * n;
  const int nblk = (int)((total + 255) / 256);
  prof_begin(4, la->stream);
  if (la->is_real) k_la_finish<1, 1><<<nblk, 256, 0, la->stream>>>(la->part, g.Mp, n, n, g.ksplit, (const double2*)c, (const double2*)c, ldc, nullptr, nullptr, 0, (double2*)la->S, n);
  else k_la_finish<0, 1><<<nblk, 256, 0, la->stream>>>(la->part, g.Mp, n, n, g.ksplit, (const double2*)c, (const double2*)c, ldc, nullptr, nullptr, 0, (double2*)la->S, n);
  LA_LAUNCH_CHECK(la);
  if ((rc = la_factor(la, n))) return rc;
  if (la->is_real) k_gram_operand<1><<<nblk, 256, 0, la->stream>>>((const double2*)la->X, n, la->fs, g.FP);
  else k_gram_operand<0><<<nblk, 256, 0, la->stream>>>((const double2*)la->X, n, la->fs, g.FP);
  prof_end(la->stream);
  LA_LAUNCH_CHECK(la);
  int h_info = 0;
  QB_CUDA(cudaMemcpyAsync(&h_info, la->info_dev, sizeof(int), cudaMemcpyDeviceToHost, la->stream));
  QB_CUDA(cudaStreamSynchronize(la->stream));
  if (info) *info = h_info;
  if (h_info != 0) { set_error("qb200_gram: overlap matrix not positive definite (potrf info > 0)"); return QB200_EINVAL; }
  // C <- C L^-H from the packed copy, chunk by chunk (a chunk's rows of C are overwritten only after they were packed)
  for (int i = 0; i < g.nchunks; i++) {
    const int ch = g.nchunks - 1 - i;
    const int gbeg = ch * g.gchunk, gcount = std::min(g.gchunk, ngw - gbeg), gpad = (gcount + 15) / 16 * 16;
    if (i > 0 && (rc = la_pack(la, g, c, ldc, gbeg, gcount, gpad))) return rc;
    if ((rc = la_back(la, g, la->W, gbeg, gcount, c, ldc, n, 1, 1))) return rc;       // L^-H is upper triangular
  }
  return QB200_OK;
}

extern "C" int qb200_residual(qb200_la* la, int ldc, int nall, const double* c, int nst, double* hc, double* a)
{
  if (!la || !c || !hc || nall < 0 || nst < 0 || ldc < la->ngw) { set_error("qb200_residual: bad argument"); return QB200_EINVAL; }
  if (nall == 0 || nst == 0) return QB200_OK;
  QB_CUDA(cudaSetDevice(la->device));
  int rc;
  const size_t cb = 2 * (size_t)ldc * nall, xb = 2 * (size_t)ldc * nst, ab = (la->is_real ? 1 : 2) * (size_t)nall * nst;
  const double* cd = c; double* xd = hc; double* ad = a;
  if (!is_device_ptr(c)) {
    if ((rc = nl_ensure(&la->st_c, &la->st_c_cap, cb))) return rc;
    QB_CUDA(cudaMemcpyAsync(la->st_c, c, cb * sizeof(double), cudaMemcpyHostToDevice, la->stream));
    cd = la->st_c;
  }
  if (!is_device_ptr(hc)) {
    if ((rc = nl_ensure(&la->st_x, &la->st_x_cap, xb))) return rc;
    QB_CUDA(cudaMemcpyAsync(la->st_x, hc, xb * sizeof(double), cudaMemcpyHostToDevice, la->stream));
    xd = la->st_x;
  }
  if (a && !is_device_ptr(a)) {
    if ((rc = nl_ensure(&la->st_a, &la->st_a_cap, ab))) return rc;
    ad = la->st_a;
  }
  if ((rc = la_residual_dev(la, ldc, nall, cd, nst, xd, ad))) return rc;
  if (xd != hc) QB_CUDA(cudaMemcpyAsync(hc, xd, xb * sizeof(double), cudaMemcpyDeviceToHost, la->stream));
  if (a && ad != a) QB_CUDA(cudaMemcpyAsync(a, ad, ab * sizeof(double), cudaMemcpyDeviceToHost, la->stream));
  if (xd != hc || (a && ad != a)) QB_CUDA(cudaStreamSynchronize(la->stream));
  return QB200_OK;
}

extern "C" int qb200_gram(qb200_la* la, int ldc, int nst, double* c, int* info)
{
  if (info) *info = 0;
  if (!la || !c || nst < 0 || ldc < la->ngw) { set_error("qb200_gram: bad argument"); return QB200_EINVAL; }
  if (nst == 0) return QB200_OK;
  QB_CUDA(cudaSetDevice(la->device));
  int rc;
  const size_t cb = 2 * (size_t)ldc * nst;
  double* cd = c;
  if (!is_device_ptr(c)) {
    if ((rc = nl_ensure(&la->st_c, &la->st_c_cap, cb))) return rc;
    QB_CUDA(cudaMemcpyAsync(la->st_c, c, cb * sizeof(double), cudaMemcpyHostToDevice, la->stream));
    cd = la->st_c;
  }
  if ((rc = la_gram_dev(la, ldc, nst, cd, info))) return rc;
  if (cd != c) {
    QB_CUDA(cudaMemcpyAsync(c, cd, cb * sizeof(double), cudaMemcpyDeviceToHost, la->stream));
    QB_CUDA(cudaStreamSynchronize(la->stream));
  }
  return QB200_OK;
}

// ------------------------------------------------------------------------------------------------ band-sharded gram
// SlaterDet::gram with the states distributed over process columns (the reference's pzherk / pzpotrf / pztrsm on the
// square context, SlaterDet.cc:1043-1143), restated for band sharding with nprow = 1:
//   (1) overlap : this rank's columns S[:, first .. first+nst) = c_all^H c_local        (1/P of the herk)
//   (2) sum of the column blocks over the ranks (every other entry of a rank's S is zero: the sum is a gather)
//   (3) apply   : Cholesky S = L L^H and L^-1 on every rank (replicated, as cheap as on one GPU), then
//                 c_local <- c_all T[:, first .. first+nst), T = L^-H                  (1/P of the trsm)
// c_all: the gathered ldc x nall block (as for qb200_residual), device pointer; c_local: ldc x nst, device pointer, may be
// the rank's own columns inside c_all.
static int la_gram_check(qb200_la* la, const char* who, int ldc, int nall, const double* c_all, int first, int nst)
{
  if (!la || !c_all || nall < 1 || nst < 0 || first < 0 || first + nst > nall || ldc < la->ngw) { set_error(std::string(who) + ": bad argument"); return QB200_EINVAL; }
  if (!is_device_ptr(c_all)) { set_error(std::string(who) + ": device pointers only (the gathered block lives on the device)"); return QB200_EINVAL; }
  return QB200_OK;
}

extern "C" int qb200_gram_overlap(qb200_la* la, int ldc, int nall, const double* c_all, int first, int nst, double* S)
{
  int rc;
  if ((rc = la_gram_check(la, "qb200_gram_overlap", ldc, nall, c_all, first, nst))) return rc;
  if (!S || !is_device_ptr(S)) { set_error("qb200_gram_overlap: S must be a device pointer (nall x nall complex)"); return QB200_EINVAL; }
  QB_CUDA(cudaSetDevice(la->device));
  QB_CUDA(cudaMemsetAsync(S, 0, 2 * (size_t)nall * nall * sizeof(double), la->stream));
  if (nst == 0) return QB200_OK;
  const LaGeom g = la_geometry(la, nall, nst, false, ldc);
  la->nchunks_last = g.nchunks;
  if ((rc = la_prepare_W(la, g))) return rc;
  const int ncols = la->is_real ? nst : 2 * nst;
  if ((rc = nl_ensure(&la->part, &la->part_cap, (size_t)g.ksplit * ncols * g.Mp))) return rc;
  const double* x = c_all + 2 * (size_t)first * ldc;
  for (int ch = 0; ch < g.nchunks; ch++) {
    const int gbeg = ch * g.gchunk, gcount = std::min(g.gchunk, la->ngw - gbeg), gpad = (gcount + 15) / 16 * 16;
    if ((rc = la_pack(la, g, c_all, ldc, gbeg, gcount, gpad))) return rc;
    if ((rc = la_fnl(la, g, la->W, gbeg, gcount, x, ldc, nst, ch > 0))) return rc;
  }
  const size_t total = (size_t)nst * nall;
  const int nblk = (int)((total + 255) / 256);
  double2* Sc = (double2*)S + (size_t)first * nall;
  if (la->is_real) k_la_finish<1, 1><<<nblk, 256, 0, la->stream>>>(la->part, g.Mp, nall, nst, g.ksplit, (const double2*)c_all, (const double2*)x, ldc, nullptr, nullptr, 0, Sc, nall);
  else k_la_finish<0, 1><<<nblk, 256, 0, la->stream>>>(la->part, g.Mp, nall, nst, g.ksplit, (const double2*)c_all, (const double2*)x, ldc, nullptr, nullptr, 0, Sc, nall);
  LA_LAUNCH_CHECK(la);
  return QB200_OK;
}

extern "C" int qb200_gram_apply(qb200_la* la, int ldc, int nall, const double* c_all, const double* S, int first, int nst, double* c_local, int* info)
{
  if (info) *info = 0;
  int rc;
  if ((rc = la_gram_check(la, "qb200_gram_apply", ldc, nall, c_all, first, nst))) return rc;
  if (!S || !is_device_ptr(S) || (nst > 0 && (!c_local || !is_device_ptr(c_local)))) { set_error("qb200_gram_apply: device pointers only"); return QB200_EINVAL; }
  QB_CUDA(cudaSetDevice(la->device));
  const int n = nall, nblkd = (n + LA_NB - 1) / LA_NB;
  const LaGeom g = la_geometry(la, nall, std::max(nst, 1), false, ldc);
  la->nchunks_last = g.nchunks;
  if ((rc = la_prepare_W(la, g))) return rc;
  if ((rc = nl_ensure(&la->fs, &la->fs_cap, (size_t)std::max(nst, 1) * g.FP))) return rc;
  if ((rc = nl_ensure(&la->S, &la->S_cap, 2 * (size_t)n * n))) return rc;
  if ((rc = nl_ensure(&la->X, &la->X_cap, 2 * (size_t)n * n))) return rc;
  if ((rc = nl_ensure(&la->Dinv, &la->Dinv_cap, 2 * (size_t)nblkd * LA_NB * LA_NB))) return rc;
  QB_CUDA(cudaMemcpyAsync(la->S, S, 2 * (size_t)n * n * sizeof(double), cudaMemcpyDeviceToDevice, la->stream));   // the factorisation is in place
  QB_CUDA(cudaMemsetAsync(la->fs, 0, (size_t)std::max(nst, 1) * g.FP * sizeof(double), la->stream));
  QB_CUDA(cudaMemsetAsync(la->X, 0, 2 * (size_t)n * n * sizeof(double), la->stream));
  QB_CUDA(cudaMemsetAsync(la->info_dev, 0, sizeof(int), la->stream));
  if ((rc = la_factor(la, n))) return rc;
  if (nst > 0) {
    const int nblk = (int)(((size_t)nst * n + 255) / 256);
    if (la->is_real) k_gram_operand_cols<1><<<nblk, 256, 0, la->stream>>>((const double2*)la->X, n, first, nst, la->fs, g.FP);
    else k_gram_operand_cols<0><<<nblk, 256, 0, la->stream>>>((const double2*)la->X, n, first, nst, la->fs, g.FP);
    LA_LAUNCH_CHECK(la);
  }
  int h_info = 0;
  QB_CUDA(cudaMemcpyAsync(&h_info, la->info_dev, sizeof(int), cudaMemcpyDeviceToHost, la->stream));
  QB_CUDA(cudaStreamSynchronize(la->stream));
  if (info) *info = h_info;
  if (h_info != 0) { set_error("qb200_gram_apply: overlap matrix not positive definite (potrf info > 0)"); return QB200_EINVAL; }
  for (int i = 0; i < g.nchunks && nst > 0; i++) {
    const int ch = g.nchunks - 1 - i;
    const int gbeg = ch * g.gchunk, gcount = std::min(g.gchunk, la->ngw - gbeg), gpad = (gcount + 15) / 16 * 16;
    if ((rc = la_pack(la, g, c_all, ldc, gbeg, gcount, gpad))) return rc;     // (a chunk's rows are packed before c_local, which may alias them, is written)
    if ((rc = la_back(la, g, la->W, gbeg, gcount, c_local, ldc, nst, 1))) return rc;
  }
  return QB200_OK;
}

extern "C" int qb200_gram_sharded(qb200_la* la, qb200_comm* comm, int ldc, int nall, const double* c_all, int first, int nst, double* c_local, int* info)
{
  if (info) *info = 0;
  int rc;
  if ((rc = la_gram_check(la, "qb200_gram_sharded", ldc, nall, c_all, first, nst))) return rc;
  if (!comm && nst != nall) { set_error("qb200_gram_sharded: a communicator is needed when the rank does not hold every state"); return QB200_EINVAL; }
  QB_CUDA(cudaSetDevice(la->device));
  if ((rc = nl_ensure(&la->Ssh, &la->Ssh_cap, 2 * (size_t)nall * nall))) return rc;
  if ((rc = qb200_gram_overlap(la, ldc, nall, c_all, first, nst, la->Ssh))) return rc;
  if (comm && (rc = qb200_allreduce_rho(comm, la->Ssh, 2ll * nall * nall, (void*)la->stream))) return rc;
  return qb200_gram_apply(la, ldc, nall, c_all, la->Ssh, first, nst, c_local, info);
}

// ------------------------------------------------------------------------------------------------ PSDA wavefunction update
// The rest of PSDAWavefunctionStepper::update after the descent direction (PSDAWavefunctionStepper.cc:93-225 real basis,
// :281-395 complex) with Preconditioner::apply(sd, ispin, ikp, -1.0) (Preconditioner.cc:118-139), so that between two H psi
// evaluations the wavefunction block never leaves the device:
//   dc[ig,n] *= -precdiag[ig]                              for ig < ngw            (padding rows untouched)
//   extrapolate:  a = sum_n occ_n sum_i f (f - f_last),  b = sum_n occ_n sum_i (f - f_last)^2   over the 2*ldc doubles of
//                 every local column (real basis: both doubled and the G = 0 row subtracted once, :146-173),
//                 summed over the ranks (dsum over the sd context, :177 / :343);  theta = -a/b, theta < -1 -> 0, min(2, theta)
//   c <- c + theta (c - c_last) + f + theta (f - f_last);  c_last <- old c;  dc_last <- f         (:192-209 / :362-379)
//   no extrapolation (first call after a reset): c <- c + f, c_last <- old c, dc_last <- f          (:211-221)
// Partial sums: one (a, b) pair per CTA in a fixed layout, reduced by one warp in index order: deterministic.
__global__ void __launch_bounds__(256) k_psda_prec_dots(double2* __restrict__ dc, const double2* __restrict__ dc_last, size_t ldc, int ngw,
                                                        const double* __restrict__ precdiag, const double* __restrict__ occ,
                                                        int is_real, int want_dots, double* __restrict__ part)
{
  // grid (ceil(ldc/256), nst): one state per blockIdx.y
  __shared__ double ra[256], rb[256];
  const int n = blockIdx.y;
  const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
  double a = 0.0, b = 0.0;
  if (i < ldc) {
    double2 f = dc[(size_t)n * ldc + i];
    if (i < (size_t)ngw) {
      const double k = -precdiag[i];
      f.x *= k; f.y *= k;
      dc[(size_t)n * ldc + i] = f;
    }
    if (want_dots) {
      const double2 l = dc_last[(size_t)n * ldc + i];
      const double dx = f.x - l.x, dy = f.y - l.y;
      double w = occ[n];
      if (is_real) w *= (i == 0 ? 1.0 : 2.0);            // G and -G; the G = 0 row counted once (:146-173)
      a = w * (f.x * dx + f.y * dy);
      b = w * (dx * dx + dy * dy);
    }
  }
  if (!want_dots) return;
  ra[threadIdx.x] = a; rb[threadIdx.x] = b;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) { ra[threadIdx.x] += ra[threadIdx.x + s]; rb[threadIdx.x] += rb[threadIdx.x + s]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const size_t blk = (size_t)n * gridDim.x + blockIdx.x;
    part[2 * blk] = ra[0]; part[2 * blk + 1] = rb[0];
  }
}
__global__ void __launch_bounds__(1024) k_psda_sum_ab(const double* __restrict__ part, size_t nblk, double* __restrict__ ab)
{
  __shared__ double ra[1024], rb[1024];
  double a = 0.0, b = 0.0;
  for (size_t i = threadIdx.x; i < nblk; i += 1024) { a += part[2 * i]; b += part[2 * i + 1]; }
  ra[threadIdx.x] = a; rb[threadIdx.x] = b;
  __syncthreads();
  for (int s = 512; s > 0; s >>= 1) {
    if (threadIdx.x < s) { ra[threadIdx.x] += ra[threadIdx.x + s]; rb[threadIdx.x] += rb[threadIdx.x + s]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) { ab[0] = ra[0]; ab[1] = rb[0]; }
}
__global__ void __launch_bounds__(256) k_psda_apply(size_t n2, double theta, double* __restrict__ c, const double* __restrict__ dc,
                                                    double* __restrict__ c_last, double* __restrict__ dc_last, int extrapolate)
{
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n2; i += (size_t)gridDim.x * blockDim.x) {
    const double x = c[i], f = dc[i];
    double xn;
    if (extrapolate) {
      const double xbar = x + theta * (x - c_last[i]);
      const double fbar = f + theta * (f - dc_last[i]);
      xn = xbar + fbar;
    } else {
      xn = x + f;
    }
    c[i] = xn; c_last[i] = x; dc_last[i] = f;
  }
}

extern "C" int qb200_psda_update(qb200_la* la, qb200_comm* comm, int ldc, int nst, double* c, double* dc, double* c_last,
                                 double* dc_last, const double* occ, const double* precdiag, int extrapolate, double* theta_out)
{
  if (!la || !c || !dc || !c_last || !dc_last || !occ || !precdiag || nst < 0 || ldc < la->ngw) { set_error("qb200_psda_update: bad argument"); return QB200_EINVAL; }
  if (theta_out) *theta_out = 0.0;
  if (nst == 0 && !comm) return QB200_OK;
  if (nst > 0 && (!is_device_ptr(c) || !is_device_ptr(dc) || !is_device_ptr(c_last) || !is_device_ptr(dc_last))) {
    set_error("qb200_psda_update: the four blocks must be device-resident (this call exists to keep the stepper on the device)");
    return QB200_EINVAL;
  }
  QB_CUDA(cudaSetDevice(la->device));
  int rc;
  const int gx = (ldc + 255) / 256;
  const size_t nblk = (size_t)gx * nst;
  // work: part[2*nblk] | ab[2] | occ[nst] | precdiag[ngw] (if host)
  const bool hp = !is_device_ptr(precdiag);
  if ((rc = nl_ensure(&la->part, &la->part_cap, 2 * nblk + 2 + nst + (hp ? la->ngw : 0) + 8))) return rc;
  double* part = la->part;
  double* ab = part + 2 * nblk;
  double* occd = ab + 2;
  const double* pd = precdiag;
  QB_CUDA(cudaMemcpyAsync(occd, occ, (size_t)nst * sizeof(double), cudaMemcpyDefault, la->stream));
  if (hp) {
    double* q = occd + nst;
    QB_CUDA(cudaMemcpyAsync(q, precdiag, (size_t)la->ngw * sizeof(double), cudaMemcpyHostToDevice, la->stream));
    pd = q;
  }
  double h_ab[2] = { 0.0, 0.0 };
  if (nst > 0) {
    k_psda_prec_dots<<<dim3(gx, nst), 256, 0, la->stream>>>((double2*)dc, (const double2*)dc_last, (size_t)ldc, la->ngw, pd, occd, la->is_real,
                                                            extrapolate ? 1 : 0, part);
    LA_LAUNCH_CHECK(la);
    if (extrapolate) {
      k_psda_sum_ab<<<1, 1024, 0, la->stream>>>(part, nblk, ab);
      LA_LAUNCH_CHECK(la);
      QB_CUDA(cudaMemcpyAsync(h_ab, ab, 2 * sizeof(double), cudaMemcpyDeviceToHost, la->stream));
      QB_CUDA(cudaStreamSynchronize(la->stream));
    }
  }
  double theta = 0.0;
  if (extrapolate) {
    if (comm && (rc = qb200_allreduce_scalars(comm, h_ab, 2))) return rc;          // dsum over the sd context (:177, :343)
    if (h_ab[1] != 0.0) theta = -h_ab[0] / h_ab[1];
    if (theta_out) *theta_out = theta;                                             // the value the reference prints before clipping
    if (theta < -1.0) theta = 0.0;
    theta = std::min(2.0, theta);
  }
  if (nst > 0) {
    k_psda_apply<<<148 * 8, 256, 0, la->stream>>>(2 * (size_t)ldc * nst, theta, c, dc, c_last, dc_last, extrapolate ? 1 : 0);
    LA_LAUNCH_CHECK(la);
  }
  return QB200_OK;
}

// ------------------------------------------------------------------------------------------------ subspace diagonalisation
// Wavefunction::diag (/root/reference/src/qball/Wavefunction.cc:1510-1715), norm-conserving, one (spin, k-point):
//   h = c^H (H c)                 (complex: h.gemm('c','n',1.0,c,cp,0.0) :1641; real: gemm('t','n',2.0) + ger(-1.0) :1538-1539)
//   w = eigenvalues of h, ascending, from its LOWER triangle (syevd / heevd 'l', :1604, :1682)
//   eigvec: z = eigenvectors, c <- c z (:1606-1609, :1684-1688)
// The reference calls LAPACK; here the n x n Hermitian problem is solved on the device by a parallel cyclic Jacobi method
// (round-robin ordering: n/2 disjoint rotations per step, n - 1 steps per sweep, quadratic convergence, every rotation unitary
// to rounding -- eigenvalues accurate to a few ulp of ||h||), and c z reuses the FP64 tensor-core GEMM of qb200_gram.
// Eigenvectors of degenerate eigenvalues and the phase of every eigenvector are as arbitrary as LAPACK's.
namespace qb200 {

// A (ne x ne, column-major) <- Hermitian matrix defined by the lower triangle of S (n x n); padding row / column ne - 1 zero
__global__ void __launch_bounds__(256) k_jac_init(const double2* __restrict__ S, int n, int ne, double2* __restrict__ A, double2* __restrict__ Z)
{
  const size_t idx = blockIdx.x * (size_t)256 + threadIdx.x;
  if (idx >= (size_t)ne * ne) return;
  const int col = (int)(idx / ne), row = (int)(idx % ne);
  double2 v = make_double2(0.0, 0.0);
  if (row < n && col < n) {
    if (row > col) v = S[(size_t)col * n + row];
    else if (row < col) { const double2 t = S[(size_t)row * n + col]; v = make_double2(t.x, -t.y); }
    else v = make_double2(S[(size_t)col * n + row].x, 0.0);
  }
  A[idx] = v;
  Z[idx] = make_double2(row == col ? 1.0 : 0.0, 0.0);
}

// round-robin pairing of step `step` (0 .. ne-2): player 0 fixed, the others rotate
__device__ __forceinline__ void jac_pair(int k, int step, int ne, int& p, int& q)
{
  const int m = ne - 1;
  int a = (k == 0) ? m : ((step + k) % m);
  int b = (step + m - k) % m;
  p = min(a, b); q = max(a, b);
}

struct JacRot { double c, s, ex, ey; };     // J[p,p] = J[q,q] = c, J[p,q] = s e^{i phi}, J[q,p] = -s e^{-i phi}; (ex, ey) = e^{i phi}

// one CTA per pair: rotation from (A[p,p], A[q,q], A[p,q]), then the COLUMN operation A <- A J, Z <- Z J on columns p, q
__global__ void __launch_bounds__(256) k_jac_cols(double2* __restrict__ A, double2* __restrict__ Z, int ne, int step, JacRot* __restrict__ rot)
{
  int p, q;
  jac_pair(blockIdx.x, step, ne, p, q);
  __shared__ JacRot R;
  if (threadIdx.x == 0) {
    const double app = A[(size_t)p * ne + p].x, aqq = A[(size_t)q * ne + q].x;
    const double2 apq = A[(size_t)q * ne + p];                    // element (row p, column q)
    const double mod = hypot(apq.x, apq.y);
    JacRot r;
    if (mod <= 1e-300 || mod * mod <= 1e-34 * fabs(app * aqq)) { r.c = 1.0; r.s = 0.0; r.ex = 1.0; r.ey = 0.0; }
    else {
      const double tau = (aqq - app) / (2.0 * mod);
      const double t = (tau >= 0.0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
      r.c = 1.0 / sqrt(1.0 + t * t); r.s = t * r.c; r.ex = apq.x / mod; r.ey = apq.y / mod;
    }
    R = r; rot[blockIdx.x] = r;
  }
  __syncthreads();
  const double c = R.c, s = R.s, ex = R.ex, ey = R.ey;
  if (s == 0.0) return;
  for (int i = threadIdx.x; i < 2 * ne; i += 256) {
    double2* M = i < ne ? A : Z;
    const int r = i < ne ? i : i - ne;
    const double2 xp = M[(size_t)p * ne + r], xq = M[(size_t)q * ne + r];
    // new_p = c xp - s e^{-i phi} xq ; new_q = s e^{i phi} xp + c xq
    const double2 eq = make_double2(ex * xq.x + ey * xq.y, ex * xq.y - ey * xq.x);      // e^{-i phi} xq
    const double2 ep = make_double2(ex * xp.x - ey * xp.y, ex * xp.y + ey * xp.x);      // e^{+i phi} xp
    M[(size_t)p * ne + r] = make_double2(c * xp.x - s * eq.x, c * xp.y - s * eq.y);
    M[(size_t)q * ne + r] = make_double2(s * ep.x + c * xq.x, s * ep.y + c * xq.y);
  }
}
// ROW operation A <- J^H A for all pairs: thread (column r, pair k)
__global__ void __launch_bounds__(256) k_jac_rows(double2* __restrict__ A, int ne, int step, const JacRot* __restrict__ rot)
{
  const int r = blockIdx.x * 256 + threadIdx.x;
  if (r >= ne) return;
  int p, q;
  jac_pair(blockIdx.y, step, ne, p, q);
  const JacRot R = rot[blockIdx.y];
  if (R.s == 0.0) return;
  const double c = R.c, s = R.s, ex = R.ex, ey = R.ey;
  const double2 xp = A[(size_t)r * ne + p], xq = A[(size_t)r * ne + q];
  // new_p = c xp - s e^{i phi} xq ; new_q = s e^{-i phi} xp + c xq
  const double2 eq = make_double2(ex * xq.x - ey * xq.y, ex * xq.y + ey * xq.x);
  const double2 ep = make_double2(ex * xp.x + ey * xp.y, ex * xp.y - ey * xp.x);
  double2 np_ = make_double2(c * xp.x - s * eq.x, c * xp.y - s * eq.y);
  double2 nq_ = make_double2(s * ep.x + c * xq.x, s * ep.y + c * xq.y);
  if (r == p) { np_.y = 0.0; }                   // the diagonal stays real, the annihilated pair exactly zero
  if (r == q) { nq_.y = 0.0; np_ = make_double2(0.0, 0.0); }
  if (r == p) nq_ = make_double2(0.0, 0.0);
  A[(size_t)r * ne + p] = np_;
  A[(size_t)r * ne + q] = nq_;
}
// sums[0] = sum_{i != j} |a_ij|^2, sums[1] = sum |a_ij|^2: every CTA reduces a fixed slice to part[2 b], part[2 b + 1]; the last
// stage (one CTA) adds the partials in index order -- deterministic
__global__ void __launch_bounds__(1024) k_jac_off(const double2* __restrict__ A, int ne, double* __restrict__ part)
{
  __shared__ double ro[1024], rt[1024];
  double o = 0.0, t = 0.0;
  const size_t total = (size_t)ne * ne, per = (total + gridDim.x - 1) / gridDim.x;
  const size_t i0 = blockIdx.x * per, i1 = i0 + per < total ? i0 + per : total;
  for (size_t i = i0 + threadIdx.x; i < i1; i += 1024) {
    const double2 a = A[i];
    const double v = a.x * a.x + a.y * a.y;
    t += v;
    if ((int)(i / ne) != (int)(i % ne)) o += v;
  }
  ro[threadIdx.x] = o; rt[threadIdx.x] = t;
  __syncthreads();
  for (int s = 512; s > 0; s >>= 1) {
    if (threadIdx.x < s) { ro[threadIdx.x] += ro[threadIdx.x + s]; rt[threadIdx.x] += rt[threadIdx.x + s]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) { part[2 * blockIdx.x] = ro[0]; part[2 * blockIdx.x + 1] = rt[0]; }
}
__global__ void k_jac_off_sum(const double* __restrict__ part, int nb, double* __restrict__ sums)
{
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    double o = 0.0, t = 0.0;
    for (int b = 0; b < nb; b++) { o += part[2 * b]; t += part[2 * b + 1]; }
    sums[0] = o; sums[1] = t;
  }
}
__global__ void __launch_bounds__(256) k_jac_diag(const double2* __restrict__ A, int ne, int n, double* __restrict__ w)
{
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i < n) w[i] = A[(size_t)i * ne + i].x;
}
// ---- blocked Jacobi: the matrix in blocks of BJ_NB; a step pairs the blocks round-robin, one CTA diagonalises each 2 BJ_NB x
// 2 BJ_NB pivot [A_PP A_PQ; A_QP A_QQ] completely in shared memory (cyclic Jacobi, no launches) and hands back its unitary U; two
// small GEMM kernels then apply U to the block columns of A and Z and U^H to the block rows of A.  A sweep is 3 (nblk - 1) launches
// instead of 2 (n - 1), and the sequential chain of rotation steps runs at shared-memory latency instead of launch latency.
constexpr int BJ_NB = 32, BJ_M = 2 * BJ_NB, BJ_LD = BJ_M + 1, BJ_PT = 1024;
static_assert(BJ_PT == 32 * BJ_NB, "one warp per pair of a pivot step");
constexpr size_t BJ_PIVOT_SMEM = 2 * (size_t)BJ_M * BJ_LD * sizeof(double2) + BJ_NB * sizeof(JacRot) + 2 * BJ_PT * sizeof(double);
constexpr size_t BJ_APPLY_SMEM = 2 * (size_t)BJ_M * BJ_LD * sizeof(double2);

__device__ __forceinline__ int bj_global(int k, int P, int Q) { return k < BJ_NB ? P * BJ_NB + k : Q * BJ_NB + (k - BJ_NB); }

// grid (nblk/2), BJ_PT threads.  U[pair][col j][row k] (BJ_M x BJ_M, column-major, dense).  inner_max: cyclic sweeps over the pivot
// (the outer sweeps finish what a pivot leaves: 3 keep the quadratic convergence, a fully converged pivot costs 4x the time)
__global__ void __launch_bounds__(BJ_PT) k_bj_pivot(const double2* __restrict__ A, int ne, int nblk, int step, double2* __restrict__ Ubuf,
                                                    double2* __restrict__ Mbuf, int inner_max)
{
  extern __shared__ __align__(16) unsigned char bj_raw[];
  double2* M = reinterpret_cast<double2*>(bj_raw);          // M[col * BJ_LD + row]
  double2* U = M + BJ_M * BJ_LD;
  JacRot* rot = reinterpret_cast<JacRot*>(U + BJ_M * BJ_LD);
  double* red = reinterpret_cast<double*>(rot + BJ_NB);      // [2][BJ_PT]
  int P, Q;
  jac_pair(blockIdx.x, step, nblk, P, Q);
  const int tid = threadIdx.x;
  for (int e = tid; e < BJ_M * BJ_M; e += BJ_PT) {
    const int col = e / BJ_M, row = e % BJ_M;
    const int gr = bj_global(row, P, Q), gc = bj_global(col, P, Q);
    double2 v;
    if (row > col) v = A[(size_t)gc * ne + gr];                                  // the lower triangle defines the Hermitian block
    else if (row < col) { const double2 t = A[(size_t)gr * ne + gc]; v = make_double2(t.x, -t.y); }
    else v = make_double2(A[(size_t)gc * ne + gr].x, 0.0);
    M[col * BJ_LD + row] = v;
    U[col * BJ_LD + row] = make_double2(row == col ? 1.0 : 0.0, 0.0);
  }
  __syncthreads();
  for (int sw = 0; sw < inner_max; sw++) {
    // off-diagonal and total norm of the pivot (fixed order)
    double o = 0.0, t = 0.0;
    for (int e = tid; e < BJ_M * BJ_M; e += BJ_PT) {
      const int col = e / BJ_M, row = e % BJ_M;
      const double2 a = M[col * BJ_LD + row];
      const double v = a.x * a.x + a.y * a.y;
      t += v;
      if (row != col) o += v;
    }
    red[tid] = o; red[BJ_PT + tid] = t;
    __syncthreads();
    for (int k = BJ_PT / 2; k > 0; k >>= 1) {
      if (tid < k) { red[tid] += red[tid + k]; red[BJ_PT + tid] += red[BJ_PT + tid + k]; }
      __syncthreads();
    }
    const bool done = !(red[0] > 1e-32 * red[BJ_PT]);
    __syncthreads();
    if (done) break;
    // one warp per pair of the step (32 warps, 32 pairs): every lane derives the rotation from the pair's own three elements
    // (broadcast reads), the warp applies it to its two columns of M and U, and after the barrier to its two rows of M -- the
    // rotation stays in registers, two barriers per step
    const int wp = tid >> 5, lane = tid & 31;
    for (int st = 0; st < BJ_M - 1; st++) {
      // the 32 rotations of the step by ONE warp (lane = pair): 32 warps deriving them redundantly would put ~300 dependent FP64
      // instructions per warp on the FP64 pipe and make the step pipe-bound
      if (wp == 0) {
        int p0, q0;
        jac_pair(lane, st, BJ_M, p0, q0);
        const double app = M[p0 * BJ_LD + p0].x, aqq = M[q0 * BJ_LD + q0].x;
        const double2 apq = M[q0 * BJ_LD + p0];               // element (row p, column q)
        const double mod = sqrt(apq.x * apq.x + apq.y * apq.y);
        JacRot r;
        r.c = 1.0; r.s = 0.0; r.ex = 1.0; r.ey = 0.0;
        if (!(mod <= 1e-150 || mod * mod <= 1e-34 * fabs(app * aqq))) {
          const double imod = 1.0 / mod;
          const double tau = 0.5 * (aqq - app) * imod;
          const double tt = (tau >= 0.0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
          r.c = rsqrt(1.0 + tt * tt); r.s = tt * r.c; r.ex = apq.x * imod; r.ey = apq.y * imod;
        }
        rot[lane] = r;
      }
      __syncthreads();
      int p, q;
      jac_pair(wp, st, BJ_M, p, q);
      const JacRot R = rot[wp];
      const double rc_ = R.c, rs = R.s, ex = R.ex, ey = R.ey;
      if (rs != 0.0) {
#pragma unroll
        for (int i = lane; i < 2 * BJ_M; i += 32) {           // column operation: columns p, q of M and of U
          double2* X = i < BJ_M ? M : U;
          const int r = i < BJ_M ? i : i - BJ_M;
          const double2 xp = X[p * BJ_LD + r], xq = X[q * BJ_LD + r];
          const double2 eq = make_double2(ex * xq.x + ey * xq.y, ex * xq.y - ey * xq.x);      // e^{-i phi} xq
          const double2 ep = make_double2(ex * xp.x - ey * xp.y, ex * xp.y + ey * xp.x);      // e^{+i phi} xp
          X[p * BJ_LD + r] = make_double2(rc_ * xp.x - rs * eq.x, rc_ * xp.y - rs * eq.y);
          X[q * BJ_LD + r] = make_double2(rs * ep.x + rc_ * xq.x, rs * ep.y + rc_ * xq.y);
        }
      }
      __syncthreads();
      if (rs != 0.0) {
#pragma unroll
        for (int r = lane; r < BJ_M; r += 32) {               // row operation: rows p, q of M
          const double2 xp = M[r * BJ_LD + p], xq = M[r * BJ_LD + q];
          const double2 eq = make_double2(ex * xq.x - ey * xq.y, ex * xq.y + ey * xq.x);
          const double2 ep = make_double2(ex * xp.x + ey * xp.y, ex * xp.y - ey * xp.x);
          double2 np_ = make_double2(rc_ * xp.x - rs * eq.x, rc_ * xp.y - rs * eq.y);
          double2 nq_ = make_double2(rs * ep.x + rc_ * xq.x, rs * ep.y + rc_ * xq.y);
          if (r == p) { np_.y = 0.0; nq_ = make_double2(0.0, 0.0); }
          if (r == q) { nq_.y = 0.0; np_ = make_double2(0.0, 0.0); }
          M[r * BJ_LD + p] = np_;
          M[r * BJ_LD + q] = nq_;
        }
      }
      __syncthreads();
    }
  }
  double2* Uo = Ubuf + (size_t)blockIdx.x * BJ_M * BJ_M;
  double2* Mo = Mbuf + (size_t)blockIdx.x * BJ_M * BJ_M;      // the diagonalised pivot: real diagonal, exact zeros where annihilated
  for (int e = tid; e < BJ_M * BJ_M; e += BJ_PT) {
    Uo[e] = U[(e / BJ_M) * BJ_LD + (e % BJ_M)];
    Mo[e] = M[(e / BJ_M) * BJ_LD + (e % BJ_M)];
  }
}

// ROWOP == 0: X[:, PQ] <- X[:, PQ] U for a tile of 64 rows (X = A, and Z when blockIdx.z == 1)
// ROWOP == 1: A[PQ, :] <- U^H A[PQ, :] for a tile of 64 columns
// grid (nblk/2, ne/64, ROWOP ? 1 : 2), 256 threads: thread -> 4 x 4 outputs
template <int ROWOP>
__global__ void __launch_bounds__(256) k_bj_apply(double2* __restrict__ A, double2* __restrict__ Z, int ne, int nblk, int step,
                                                  const double2* __restrict__ Ubuf, const double2* __restrict__ Mbuf)
{
  extern __shared__ __align__(16) unsigned char bj_raw[];
  double2* T = reinterpret_cast<double2*>(bj_raw);          // T[k * BJ_LD + x]: x = row (column op) or column (row op) inside the tile
  double2* Cc = T + BJ_M * BJ_LD;                           // Cc[k * BJ_LD + j] = coefficient of input k in output j
  int P, Q;
  jac_pair(blockIdx.x, step, nblk, P, Q);
  double2* X = (!ROWOP && blockIdx.z == 1) ? Z : A;
  const int x0 = blockIdx.y * BJ_M, tid = threadIdx.x;
  const double2* U = Ubuf + (size_t)blockIdx.x * BJ_M * BJ_M;
  for (int e = tid; e < BJ_M * BJ_M; e += 256) {
    const int a = e / BJ_M, b = e % BJ_M;
    if (!ROWOP) {
      T[a * BJ_LD + b] = X[(size_t)bj_global(a, P, Q) * ne + x0 + b];            // k = a (pivot column), x = b (row): contiguous in b
      Cc[b * BJ_LD + a] = U[(size_t)a * BJ_M + b];                                 // U[col j = a][row k = b] -> Cc[k][j]
    } else {
      T[b * BJ_LD + a] = X[(size_t)(x0 + a) * ne + bj_global(b, P, Q)];          // x = a (column), k = b (pivot row): contiguous in b
      const double2 u = U[(size_t)a * BJ_M + b];                                   // U[col i = a][row k = b]: coefficient conj(U[k][i])
      Cc[b * BJ_LD + a] = make_double2(u.x, -u.y);
    }
  }
  __syncthreads();
  const int j0 = (tid >> 4) * 4, xx0 = (tid & 15) * 4;
  double2 acc[4][4];
#pragma unroll
  for (int j = 0; j < 4; j++)
#pragma unroll
    for (int x = 0; x < 4; x++) acc[j][x] = make_double2(0.0, 0.0);
  for (int k = 0; k < BJ_M; k++) {
    double2 t[4], c[4];
#pragma unroll
    for (int x = 0; x < 4; x++) t[x] = T[k * BJ_LD + xx0 + x];
#pragma unroll
    for (int j = 0; j < 4; j++) c[j] = Cc[k * BJ_LD + j0 + j];
#pragma unroll
    for (int j = 0; j < 4; j++)
#pragma unroll
      for (int x = 0; x < 4; x++) {
        acc[j][x].x += c[j].x * t[x].x - c[j].y * t[x].y;
        acc[j][x].y += c[j].x * t[x].y + c[j].y * t[x].x;
      }
  }
#pragma unroll
  for (int j = 0; j < 4; j++)
#pragma unroll
    for (int x = 0; x < 4; x++) {
      if (!ROWOP) X[(size_t)bj_global(j0 + j, P, Q) * ne + x0 + xx0 + x] = acc[j][x];
      else {
        // the pivot block itself takes the values the pivot kernel reached in shared memory (U^H M U up to rounding): its
        // annihilated elements are EXACTLY zero, as in the element-wise method -- without this the off-diagonal norm of a
        // large matrix stalls at the rounding noise of the block GEMMs
        const int col = x0 + xx0 + x, cb = col / BJ_NB;
        double2 v = acc[j][x];
        if (cb == P || cb == Q) v = Mbuf[(size_t)blockIdx.x * BJ_M * BJ_M + (size_t)((cb == P ? 0 : BJ_NB) + col % BJ_NB) * BJ_M + j0 + j];
        X[(size_t)col * ne + bj_global(j0 + j, P, Q)] = v;
      }
    }
}

// second operand of c z: T[m, col] = Z[m, perm[col]] in the layout la_back reads (cf. k_gram_operand)
template <int IS_REAL>
__global__ void __launch_bounds__(256) k_diag_operand(const double2* __restrict__ Z, int ne, int n, const int* __restrict__ perm, double* __restrict__ fs, int FP)
{
  const size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (idx >= (size_t)n * n) return;
  const int col = (int)(idx / n), m = (int)(idx % n);
  const double2 t = Z[(size_t)perm[col] * ne + m];
  if (IS_REAL) fs[(size_t)col * FP + m] = t.x;
  else {
    double* o = fs + (size_t)col * FP + (m >> 3) * 24 + (m & 7);
    o[0] = t.x; o[8] = t.y; o[16] = t.x + t.y;
  }
}

}  // namespace qb200

static int la_diag_dev(qb200_la* la, int ldc, int n, double* c, const double* hc, int eigvec, double* w_host, int* sweeps_out)
{
  int rc;
  const LaGeom g = la_geometry(la, n, n, false, ldc, true);   // h is read from its lower triangle (syevd / heevd 'l')
  la->nchunks_last = g.nchunks;
  if ((rc = la_prepare_W(la, g))) return rc;
  const int ncols = la->is_real ? n : 2 * n;
  // blocked Jacobi (default): the matrix padded to an even number of 32 x 32 blocks; QB200_JACOBI_BLOCK=0: the element-wise
  // parallel cyclic Jacobi (two launches per rotation step), padded to an even order
  bool blocked = true;
  if (const char* e = getenv("QB200_JACOBI_BLOCK")) blocked = e[0] != '0';
  const int bj_nblk = (((n + BJ_NB - 1) / BJ_NB) + 1) & ~1;
  const int ne = blocked ? bj_nblk * BJ_NB : (n + 1) & ~1;
  if ((rc = nl_ensure(&la->part, &la->part_cap, std::max<size_t>((size_t)g.ksplit * ncols * g.Mp, 256)))) return rc;   // (>= the off-norm partials)
  if ((rc = nl_ensure(&la->fs, &la->fs_cap, (size_t)n * g.FP))) return rc;
  if ((rc = nl_ensure(&la->S, &la->S_cap, 2 * (size_t)ne * ne + 8))) return rc;
  if ((rc = nl_ensure(&la->X, &la->X_cap, 2 * (size_t)ne * ne))) return rc;
  // Dinv doubles as the small work area: A (ne x ne) | rot[ne/2] | sums[2] | w[n] | perm[n] | U of every pivot (blocked)
  const size_t rot_d = (sizeof(JacRot) * (size_t)(ne / 2) + 7) / 8;
  const size_t small_d = (rot_d + 2 + n + (n + 1) / 2 + 8 + 1) & ~(size_t)1;
  const size_t ubuf_d = blocked ? 2 * (size_t)(bj_nblk / 2) * BJ_M * BJ_M : 0;     // U of every pivot; the same again for the diagonalised pivots
  if ((rc = nl_ensure(&la->Dinv, &la->Dinv_cap, 2 * (size_t)ne * ne + small_d + 2 * ubuf_d))) return rc;
  const int ngw = la->ngw;
  // S = c^H (H c)
  for (int ch = 0; ch < g.nchunks; ch++) {
    const int gbeg = ch * g.gchunk, gcount = std::min(g.gchunk, ngw - gbeg), gpad = (gcount + 15) / 16 * 16;
    if ((rc = la_pack(la, g, c, ldc, gbeg, gcount, gpad))) return rc;
    if ((rc = la_fnl(la, g, la->W, gbeg, gcount, hc, ldc, n, ch > 0, 1))) return rc;
  }
  const size_t total = (size_t)n * n;
  const int nblk = (int)((total + 255) / 256);
  if (la->is_real) k_la_finish<1, 1><<<nblk, 256, 0, la->stream>>>(la->part, g.Mp, n, n, g.ksplit, (const double2*)c, (const double2*)hc, ldc, nullptr, nullptr, 0, (double2*)la->S, n);
  else k_la_finish<0, 1><<<nblk, 256, 0, la->stream>>>(la->part, g.Mp, n, n, g.ksplit, (const double2*)c, (const double2*)hc, ldc, nullptr, nullptr, 0, (double2*)la->S, n);
  LA_LAUNCH_CHECK(la);
  double2* A = (double2*)la->Dinv;
  double2* Z = (double2*)la->X;
  JacRot* rot = (JacRot*)(la->Dinv + 2 * (size_t)ne * ne);
  double* sums = la->Dinv + 2 * (size_t)ne * ne + rot_d;
  double* wd = sums + 2;
  int* perm_d = (int*)(wd + n);
  k_jac_init<<<(int)(((size_t)ne * ne + 255) / 256), 256, 0, la->stream>>>((const double2*)la->S, n, ne, A, Z);
  LA_LAUNCH_CHECK(la);
  int sweeps = 0;
  const int maxsweep = 30;                                  // the reference's own jacobi() uses the same cap (Wavefunction.cc:1594)
  la->jacobi_blocked = blocked;
  double2 *Ubuf = nullptr, *Mbuf = nullptr;
  int bj_inner = 1;
  if (blocked) {
    Ubuf = (double2*)(la->Dinv + 2 * (size_t)ne * ne + small_d);
    Mbuf = Ubuf + ubuf_d / 2;
    // cyclic sweeps over a pivot per visit: ONE (measured, 768 states: 67.8 ms against 99.7 / 126.1 ms with two / three -- the
    // number of outer sweeps stays the same, 11; the element-wise method with a graph: 133 ms); a single pivot = the whole
    // matrix is converged inside the kernel
    bj_inner = bj_nblk == 2 ? 12 : 1;
    if (const char* e = getenv("QB200_BJ_INNER")) bj_inner = std::max(1, atoi(e));
    QB_CUDA(cudaFuncSetAttribute(k_bj_pivot, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BJ_PIVOT_SMEM));
    QB_CUDA(cudaFuncSetAttribute(k_bj_apply<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BJ_APPLY_SMEM));
    QB_CUDA(cudaFuncSetAttribute(k_bj_apply<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BJ_APPLY_SMEM));
  }
  const int launches_per_sweep = blocked ? 3 * (bj_nblk - 1) : 2 * (ne - 1);
  auto sweep = [&]() {
    if (blocked) {
      for (int step = 0; step < bj_nblk - 1; step++) {
        k_bj_pivot<<<bj_nblk / 2, BJ_PT, BJ_PIVOT_SMEM, la->stream>>>(A, ne, bj_nblk, step, Ubuf, Mbuf, bj_inner);
        k_bj_apply<0><<<dim3(bj_nblk / 2, ne / BJ_M, 2), 256, BJ_APPLY_SMEM, la->stream>>>(A, Z, ne, bj_nblk, step, Ubuf, Mbuf);
        k_bj_apply<1><<<dim3(bj_nblk / 2, ne / BJ_M, 1), 256, BJ_APPLY_SMEM, la->stream>>>(A, Z, ne, bj_nblk, step, Ubuf, Mbuf);
      }
    } else {
      for (int step = 0; step < ne - 1; step++) {
        k_jac_cols<<<ne / 2, 256, 0, la->stream>>>(A, Z, ne, step, rot);
        k_jac_rows<<<dim3((ne + 255) / 256, ne / 2), 256, 0, la->stream>>>(A, ne, step, rot);
      }
    }
  };
  // A sweep is a fixed sequence of small launches (same pointers, same steps every sweep): on a capturable stream it is
  // captured ONCE into a CUDA graph and replayed, which also takes the host's launch jitter out of the loop; the legacy
  // default stream cannot be captured and launches directly.  QB200_JACOBI_GRAPH=0: direct launches everywhere.
  cudaGraphExec_t gexec = nullptr;
  const int JAC_OFF_CTAS = 64;                                // (la->part, free between the two GEMMs, holds the 128 partial sums)
  {
    const char* e = getenv("QB200_JACOBI_GRAPH");
    const bool want = !(e && e[0] == '0') && launches_per_sweep > 8 && la->stream != 0 && la->stream != cudaStreamLegacy;
    if (want && cudaStreamBeginCapture(la->stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
      sweep();
      cudaGraph_t graph = nullptr;
      if (cudaStreamEndCapture(la->stream, &graph) == cudaSuccess && graph) {
        if (cudaGraphInstantiate(&gexec, graph, 0) != cudaSuccess) gexec = nullptr;
        cudaGraphDestroy(graph);
      }
      cudaGetLastError();                                   // a failed capture leaves no sticky error; the direct path takes over
    }
  }
  la->jacobi_graph = gexec != nullptr;
  for (; sweeps < maxsweep; sweeps++) {
    double h[2];
    k_jac_off<<<JAC_OFF_CTAS, 1024, 0, la->stream>>>(A, ne, la->part);
    k_jac_off_sum<<<1, 32, 0, la->stream>>>(la->part, JAC_OFF_CTAS, sums);
    LA_LAUNCH_CHECK(la);
    QB_CUDA(cudaMemcpyAsync(h, sums, sizeof h, cudaMemcpyDeviceToHost, la->stream));
    QB_CUDA(cudaStreamSynchronize(la->stream));
    if (!(h[0] > 1e-30 * h[1])) break;                      // off-diagonal norm below 1e-15 ||h||
    if (gexec) {
      const cudaError_t eg = cudaGraphLaunch(gexec, la->stream);
      if (eg != cudaSuccess) { cudaGraphExecDestroy(gexec); return qb200::cuda_fail(eg, "jacobi sweep (graph)", __FILE__, __LINE__); }
    } else {
      sweep();
    }
    la->launches += launches_per_sweep;
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { if (gexec) cudaGraphExecDestroy(gexec); return qb200::cuda_fail(e, "jacobi sweep", __FILE__, __LINE__); }
  }
  if (gexec) { cudaStreamSynchronize(la->stream); cudaGraphExecDestroy(gexec); }
  if (sweeps_out) *sweeps_out = sweeps;
  k_jac_diag<<<(n + 255) / 256, 256, 0, la->stream>>>(A, ne, n, wd);
  LA_LAUNCH_CHECK(la);
  std::vector<double> w(n);
  QB_CUDA(cudaMemcpyAsync(w.data(), wd, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, la->stream));
  QB_CUDA(cudaStreamSynchronize(la->stream));
  std::vector<int> perm(n);
  for (int i = 0; i < n; i++) perm[i] = i;
  std::stable_sort(perm.begin(), perm.end(), [&](int a, int b) { return w[a] < w[b]; });        // ascending, as LAPACK
  for (int i = 0; i < n; i++) w_host[i] = w[perm[i]];
  if (!eigvec) return QB200_OK;
  QB_CUDA(cudaMemcpyAsync(perm_d, perm.data(), (size_t)n * sizeof(int), cudaMemcpyHostToDevice, la->stream));
  if (g.FP != (g.m3 ? 3 * n : n)) QB_CUDA(cudaMemsetAsync(la->fs, 0, (size_t)n * g.FP * sizeof(double), la->stream));
  if (la->is_real) k_diag_operand<1><<<nblk, 256, 0, la->stream>>>(Z, ne, n, perm_d, la->fs, g.FP);
  else k_diag_operand<0><<<nblk, 256, 0, la->stream>>>(Z, ne, n, perm_d, la->fs, g.FP);
  LA_LAUNCH_CHECK(la);
  // c <- c z from the packed copy, chunk by chunk (a chunk's rows of c are overwritten only after they were packed)
  for (int i = 0; i < g.nchunks; i++) {
    const int ch = g.nchunks - 1 - i;
    const int gbeg = ch * g.gchunk, gcount = std::min(g.gchunk, ngw - gbeg), gpad = (gcount + 15) / 16 * 16;
    if (i > 0 && (rc = la_pack(la, g, c, ldc, gbeg, gcount, gpad))) return rc;
    if ((rc = la_back(la, g, la->W, gbeg, gcount, c, ldc, n, 1, 0))) return rc;
  }
  QB_CUDA(cudaStreamSynchronize(la->stream));     // perm (host vector) was uploaded asynchronously
  return QB200_OK;
}

extern "C" int qb200_diag(qb200_la* la, int ldc, int nst, double* c, const double* hc, int eigvec, double* w, int* sweeps)
{
  if (sweeps) *sweeps = 0;
  if (!la || !c || !hc || !w || nst < 0 || ldc < la->ngw) { set_error("qb200_diag: bad argument"); return QB200_EINVAL; }
  if (nst == 0) return QB200_OK;
  if (is_device_ptr(w)) { set_error("qb200_diag: w is a host array (valarray<double> in the reference)"); return QB200_EINVAL; }
  QB_CUDA(cudaSetDevice(la->device));
  int rc;
  const size_t cb = 2 * (size_t)ldc * nst;
  double* cd = c; const double* hd = hc;
  if (!is_device_ptr(c)) {
    if ((rc = nl_ensure(&la->st_c, &la->st_c_cap, cb))) return rc;
    QB_CUDA(cudaMemcpyAsync(la->st_c, c, cb * sizeof(double), cudaMemcpyHostToDevice, la->stream));
    cd = la->st_c;
  }
  if (!is_device_ptr(hc)) {
    if ((rc = nl_ensure(&la->st_x, &la->st_x_cap, cb))) return rc;
    QB_CUDA(cudaMemcpyAsync(la->st_x, hc, cb * sizeof(double), cudaMemcpyHostToDevice, la->stream));
    hd = la->st_x;
  }
  if ((rc = la_diag_dev(la, ldc, nst, cd, hd, eigvec, w, sweeps))) return rc;
  if (eigvec && cd != c) {
    QB_CUDA(cudaMemcpyAsync(c, cd, cb * sizeof(double), cudaMemcpyDeviceToHost, la->stream));
    QB_CUDA(cudaStreamSynchronize(la->stream));
  }
  return QB200_OK;
}
