// qball_b200/csrc/nonlocal.cu -- NonLocalPotential::energy, norm-conserving branch
// (/root/reference/src/qball/NonLocalPotential.cc:1909-2171, 2628-2643) as two fused FP64 tensor-core GEMMs.
//
// The reference materialises anl[ig,(ia,ipr)] = twnl[ipr][ig] * (-i)^l * exp(-i (k+G).tau_ia)  (:1959-2036) for blocks
// of <=128 atoms and calls BLAS:  fnl = anl^H c (:2050-2068),  E_nl += occ wt/omega |fnl|^2 (:2106-2148),
// cp += anl (wt/omega fnl) (:2150-2171).  Here anl is never stored: both GEMM kernels regenerate their anl tiles on the
// fly (one FP64 sincos per (atom, G) per tile, shared by the atom's projectors) straight into shared memory and feed
// mma.sync.m8n8k4.f64 (DMMA -- tcgen05 has no FP64 kind).  Complex arithmetic is mapped on real DMMA:
//   k_fnl : out[p, (n,re|im)] = sum_k A[p][k] B[k][(n,re|im)],  k over the 2*ngw reals, B = [c_n, J c_n]
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

#define NL_PITCH 36            // doubles per shared-memory tile row: 32 + 4 -> conflict-free DMMA fragment loads
#define NL_KSTEP 32            // reals of the reduction dimension per stage
#define NL_SMEM_BYTES (192 * NL_PITCH * 8)

__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b)
{
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

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
};

// warp-level 32x32 tile: acc[i][j] is the m8n8 tile (i,j); As rows = M index, Bs rows = N index, both k-contiguous
__device__ __forceinline__ void warp_mma_32x32(const double* As, const double* Bs, double (&acc)[4][4][2], int lane)
{
  const int r = lane >> 2, kq = lane & 3;
#pragma unroll
  for (int k4 = 0; k4 < NL_KSTEP / 4; k4++) {
    double a[4], b[4];
#pragma unroll
    for (int i = 0; i < 4; i++) a[i] = As[(i * 8 + r) * NL_PITCH + k4 * 4 + kq];
#pragma unroll
    for (int j = 0; j < 4; j++) b[j] = Bs[(j * 8 + r) * NL_PITCH + k4 * 4 + kq];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
      for (int j = 0; j < 4; j++) dmma(acc[i][j][0], acc[i][j][1], a[i], b[j]);
  }
}

// ------------------------------------------------------------------------------------------------ fnl = anl^H c
// grid (ceil(M/64), ceil(ncols/128), ksplit); block 256 (8 warps as 2 x 4 of 32x32).
// part[(ks*ncols + col)*Mp + p], ncols = IS_REAL ? nst : 2*nst, col = n or 2n+{re,im}
template <int IS_REAL>
__global__ void __launch_bounds__(256, 2) k_fnl(NlSpecies S, int ngw, const double* __restrict__ kpgx,
                                                const double2* __restrict__ c, size_t ldc, int nst, int gchunk,
                                                double* __restrict__ part, int Mp)
{
  extern __shared__ __align__(16) double nl_smem[];
  double* As = nl_smem;                      // [64][NL_PITCH]
  double* Bs = nl_smem + 64 * NL_PITCH;      // [128][NL_PITCH]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp >> 2, wn = warp & 3;
  const int p0 = blockIdx.x * 64, col0 = blockIdx.y * 128;
  const int ncols = IS_REAL ? nst : 2 * nst;
  const int g0 = blockIdx.z * gchunk, g1 = min(g0 + gchunk, ngw);
  const int ia0 = p0 / S.npr;
  const int ia1 = min((p0 + 63) / S.npr, S.na - 1);
  const int nat = ia1 - ia0 + 1;
  double acc[4][4][2];
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) acc[i][j][0] = acc[i][j][1] = 0.0;

  for (int gs = g0; gs < g1; gs += NL_KSTEP / 2) {
    __syncthreads();
    // A tile: 64 projector columns x 16 plane waves
    for (int i = tid; i < 64 * (NL_KSTEP / 2); i += blockDim.x) {       // clear (tile edges, atoms cut by the tile)
      As[(i >> 4) * NL_PITCH + 2 * (i & 15)] = 0.0;
      As[(i >> 4) * NL_PITCH + 2 * (i & 15) + 1] = 0.0;
    }
    __syncthreads();
    for (int w = tid; w < nat * (NL_KSTEP / 2); w += blockDim.x) {
      const int gl = w & 15, ia = ia0 + (w >> 4), g = gs + gl;
      if (g < g1) {
        const double arg = -(kpgx[g] * S.tau[3 * ia] + kpgx[ngw + g] * S.tau[3 * ia + 1] + kpgx[2 * (size_t)ngw + g] * S.tau[3 * ia + 2]);
        double sn, cs;
        sincos(arg, &sn, &cs);
        for (int ipr = 0; ipr < S.npr; ipr++) {
          const int pl = ia * S.npr + ipr - p0;
          if (pl >= 0 && pl < 64) {
            double2 a = anl_value(S.lproj[ipr], S.twnl[(size_t)ipr * ngw + g], sn, cs);
            if (IS_REAL && g == 0) a.x *= 0.5;       // G=0 counted once: dger fix, NonLocalPotential.cc:2078-2080
            As[pl * NL_PITCH + 2 * gl] = a.x;
            As[pl * NL_PITCH + 2 * gl + 1] = a.y;
          }
        }
      }
    }
    // B tile: 128 real columns x 16 plane waves
    if (IS_REAL) {
      for (int w = tid; w < 128 * (NL_KSTEP / 2); w += blockDim.x) {
        const int gl = w & 15, nl = w >> 4, n = col0 + nl, g = gs + gl;
        double2 v = make_double2(0.0, 0.0);
        if (n < nst && g < g1) v = c[(size_t)n * ldc + g];
        Bs[nl * NL_PITCH + 2 * gl] = v.x;
        Bs[nl * NL_PITCH + 2 * gl + 1] = v.y;
      }
    } else {
      for (int w = tid; w < 64 * (NL_KSTEP / 2); w += blockDim.x) {
        const int gl = w & 15, nl = w >> 4, n = (col0 >> 1) + nl, g = gs + gl;
        double2 v = make_double2(0.0, 0.0);
        if (n < nst && g < g1) v = c[(size_t)n * ldc + g];
        // Re fnl = sum a_re c_re + a_im c_im ; Im fnl = sum a_re c_im - a_im c_re   (conj(a) * c)
        Bs[(2 * nl) * NL_PITCH + 2 * gl] = v.x;
        Bs[(2 * nl) * NL_PITCH + 2 * gl + 1] = v.y;
        Bs[(2 * nl + 1) * NL_PITCH + 2 * gl] = v.y;
        Bs[(2 * nl + 1) * NL_PITCH + 2 * gl + 1] = -v.x;
      }
    }
    __syncthreads();
    warp_mma_32x32(As + wm * 32 * NL_PITCH, Bs + wn * 32 * NL_PITCH, acc, lane);
  }
  const int r = lane >> 2, cq = lane & 3;
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++)
#pragma unroll
      for (int e = 0; e < 2; e++) {
        const int p = p0 + wm * 32 + i * 8 + r;
        const int col = col0 + wn * 32 + j * 8 + 2 * cq + e;
        if (p < S.M && col < ncols) part[((size_t)blockIdx.z * ncols + col) * Mp + p] = acc[i][j][e];
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

// ------------------------------------------------------------------------------------------------ cp += anl * fs
// grid (ceil(ngw/64), ceil(nst/64)); block 256 (8 warps as 4 x 2 of 32x32): 128 output reals (64 G) x 64 states
template <int IS_REAL>
__global__ void __launch_bounds__(256, 2) k_back(NlSpecies S, int ngw, const double* __restrict__ kpgx,
                                                 const double* __restrict__ fs, int Mp, double2* __restrict__ cp, size_t ldc,
                                                 int nst)
{
  extern __shared__ __align__(16) double nl_smem[];
  double* As = nl_smem;                      // [128][NL_PITCH]
  double* Bs = nl_smem + 128 * NL_PITCH;     // [64][NL_PITCH]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp >> 1, wn = warp & 1;
  const int g0 = blockIdx.x * 64, n0 = blockIdx.y * 64;
  constexpr int PSTEP = IS_REAL ? NL_KSTEP : NL_KSTEP / 2;   // projector columns per stage
  double acc[4][4][2];
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) acc[i][j][0] = acc[i][j][1] = 0.0;

  for (int ps = 0; ps < S.M; ps += PSTEP) {
    __syncthreads();
    for (int i = tid; i < 128 * NL_KSTEP; i += blockDim.x) As[(i >> 5) * NL_PITCH + (i & 31)] = 0.0;
    __syncthreads();
    const int ia0 = ps / S.npr;
    const int ia1 = min((ps + PSTEP - 1) / S.npr, S.na - 1);
    const int nat = ia1 - ia0 + 1;
    for (int w = tid; w < nat * 64; w += blockDim.x) {
      const int gl = w & 63, ia = ia0 + (w >> 6), g = g0 + gl;
      if (g < ngw) {
        const double arg = -(kpgx[g] * S.tau[3 * ia] + kpgx[ngw + g] * S.tau[3 * ia + 1] + kpgx[2 * (size_t)ngw + g] * S.tau[3 * ia + 2]);
        double sn, cs;
        sincos(arg, &sn, &cs);
        for (int ipr = 0; ipr < S.npr; ipr++) {
          const int pl = ia * S.npr + ipr - ps;
          if (pl >= 0 && pl < PSTEP && ps + pl < S.M) {
            const double2 a = anl_value(S.lproj[ipr], S.twnl[(size_t)ipr * ngw + g], sn, cs);
            if (IS_REAL) {
              As[(2 * gl) * NL_PITCH + pl] = a.x;
              As[(2 * gl + 1) * NL_PITCH + pl] = a.y;
            } else {
              // (a_re + i a_im)(f_re + i f_im): rows (g,re),(g,im) ; reduction index (p,re),(p,im)
              As[(2 * gl) * NL_PITCH + 2 * pl] = a.x;
              As[(2 * gl) * NL_PITCH + 2 * pl + 1] = -a.y;
              As[(2 * gl + 1) * NL_PITCH + 2 * pl] = a.y;
              As[(2 * gl + 1) * NL_PITCH + 2 * pl + 1] = a.x;
            }
          }
        }
      }
    }
    // B tile: 64 states x NL_KSTEP reduction entries, fs is [n][p] (complex interleaved, or real at Gamma)
    for (int w = tid; w < 64 * NL_KSTEP; w += blockDim.x) {
      const int kk = w & 31, nl = w >> 5, n = n0 + nl;
      double v = 0.0;
      if (n < nst) {
        if (IS_REAL) { if (ps + kk < S.M) v = fs[(size_t)n * Mp + ps + kk]; }
        else { if (ps + (kk >> 1) < S.M) v = fs[2 * ((size_t)n * Mp + ps) + kk]; }
      }
      Bs[nl * NL_PITCH + kk] = v;
    }
    __syncthreads();
    warp_mma_32x32(As + wm * 32 * NL_PITCH, Bs + wn * 32 * NL_PITCH, acc, lane);
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
  cudaDeviceProp prop;
  QB_CUDA(cudaGetDeviceProperties(&prop, device));
  nl->nsm = prop.multiProcessorCount;
  const double* d;
  int rc = nl_upload(nl, kpgx, 3 * (size_t)ngw, &d);
  if (rc) { qb200_nl_destroy(nl); return rc; }
  nl->kpgx = const_cast<double*>(d);
  QB_CUDA(cudaMalloc((void**)&nl->enl_dev, sizeof(double)));
  QB_CUDA(cudaFuncSetAttribute(k_fnl<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, NL_SMEM_BYTES));
  QB_CUDA(cudaFuncSetAttribute(k_fnl<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, NL_SMEM_BYTES));
  QB_CUDA(cudaFuncSetAttribute(k_back<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, NL_SMEM_BYTES));
  QB_CUDA(cudaFuncSetAttribute(k_back<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, NL_SMEM_BYTES));
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
  s.lproj = nullptr; s.wt = nullptr; s.twnl = nullptr; s.tau = nullptr;
  if (s.M > 0) {
    if (!lproj || !wt || !twnl || !tau) { set_error("qb200_nl_add_species: null table"); return QB200_EINVAL; }
    for (int i = 0; i < npr; i++) if (lproj[i] < 0 || lproj[i] > 3) { set_error("qb200_nl_add_species: l > 3 unsupported (as in the reference)"); return QB200_EUNSUPPORTED; }
    int rc;
    if ((rc = nl_upload(nl, lproj, npr, &s.lproj)) || (rc = nl_upload(nl, wt, npr, &s.wt)) ||
        (rc = nl_upload(nl, twnl, (size_t)npr * nl->ngw, &s.twnl)) || (rc = nl_upload(nl, tau, 3 * (size_t)na, &s.tau))) return rc;
  }
  nl->sp.push_back(s);
  return QB200_OK;
}

extern "C" int qb200_nl_set_positions(qb200_nl* nl, int is, const double* tau)
{
  if (!nl || is < 0 || is >= (int)nl->sp.size() || !tau) { set_error("qb200_nl_set_positions: bad argument"); return QB200_EINVAL; }
  QB_CUDA(cudaSetDevice(nl->device));
  if (nl->sp[is].na > 0 && nl->sp[is].tau)
    QB_CUDA(cudaMemcpy(const_cast<double*>(nl->sp[is].tau), tau, 3 * (size_t)nl->sp[is].na * sizeof(double), cudaMemcpyHostToDevice));
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
  for (const NlSpecies& S : nl->sp) {
    if (S.M <= 0) continue;
    const int Mp = S.M;
    const int mt = (S.M + 63) / 64, nt = (ncols + 127) / 128;
    // split K so that ~4 waves of CTAs exist; chunks are multiples of 16 plane waves
    int ksplit = std::max(1, (4 * 2 * nl->nsm + mt * nt - 1) / (mt * nt));
    ksplit = std::min(ksplit, std::max(1, nl->ngw / 256));
    int gchunk = (nl->ngw + ksplit - 1) / ksplit;
    gchunk = ((gchunk + 15) / 16) * 16;
    ksplit = (nl->ngw + gchunk - 1) / gchunk;
    if ((rc = nl_ensure(&nl->part, &nl->part_cap, (size_t)ksplit * ncols * Mp))) return rc;
    if ((rc = nl_ensure(&nl->fs, &nl->fs_cap, 2 * (size_t)nst * Mp))) return rc;
    const size_t total = (size_t)nst * S.M;
    const int nblk = (int)((total + 255) / 256);
    if ((rc = nl_ensure(&nl->eblk, &nl->eblk_cap, nblk))) return rc;
    dim3 g1(mt, nt, ksplit);
    prof_begin(3, nl->stream);
    if (nl->is_real) k_fnl<1><<<g1, 256, NL_SMEM_BYTES, nl->stream>>>(S, nl->ngw, nl->kpgx, (const double2*)c, ldc, nst, gchunk, nl->part, Mp);
    else k_fnl<0><<<g1, 256, NL_SMEM_BYTES, nl->stream>>>(S, nl->ngw, nl->kpgx, (const double2*)c, ldc, nst, gchunk, nl->part, Mp);
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
      dim3 g2((nl->ngw + 63) / 64, (nst + 63) / 64);
      prof_begin(5, nl->stream);
      if (nl->is_real) k_back<1><<<g2, 256, NL_SMEM_BYTES, nl->stream>>>(S, nl->ngw, nl->kpgx, nl->fs, Mp, (double2*)cp, ldc, nst);
      else k_back<0><<<g2, 256, NL_SMEM_BYTES, nl->stream>>>(S, nl->ngw, nl->kpgx, nl->fs, Mp, (double2*)cp, ldc, nst);
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
