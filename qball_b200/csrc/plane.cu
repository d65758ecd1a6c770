// qball_b200/csrc/plane.cu -- instantiation and launch of the plane-fused xy kernels (plane_kernels.cuh).
// Own translation unit: the __noinline__ radix passes are shared by all kernels of a translation unit and compiled
// for the tightest register budget among them; here that is 65536/448 = 144 registers (128 elsewhere).
#include "plane_static.cuh"
#include "plane_tmem.cuh"
#include <cmath>
#include <algorithm>
#include <cstdlib>

namespace qb200 {

// shapes compiled in (PlaneShape<np0, np1, xsplit, xskip, ysplit, yskip, groups, threads per group>)
typedef PlaneShape<112, 112, 26, 60, 26, 60, 7, 64> ShapeMgO216;   // examples/MgO216: 112^3 grid, |h|,|k| <= 25
typedef PlaneShape<112, 112, 26, 60, 26, 60, 14, 32> ShapeMgO216w;     // one warp per 8-column block (default)
typedef PlaneShape<112, 112, 26, 60, 26, 60, 14, 32, 114> ShapeMgO216ww; // + warp-owned x phase (k_plane_w, QB200_PLANE_W=1) // same, one warp per 8-column block (QB200_GROUP_THREADS=32)

// threads per group the plan should use when nothing else is requested: the compiled MgO216 shape runs one warp per
// 8-column block (14 blocks, all in one round, warps drift out of phase: 4 % faster than 7 groups of 64)
int plane_preferred_gthreads(int np0, int np1, int ksplit, int kskip)
{
  typedef ShapeMgO216w W;
  if (const char* e = getenv("QB200_NO_STATIC")) if (e[0] == '1') return 64;
  return (np0 == W::NP0 && np1 == W::NP1 && ksplit == W::YSPLIT && kskip == W::YSKIP) ? W::GT : 64;
}

// row pitch of the shared-memory plane: np0 | 1 (odd: conflict-free whole-CTA row passes) unless the warp-owned MgO216
// geometry will run, whose 4-row passes need a pitch of 2 (mod 8) slots
int plane_preferred_pitch(int np0, int np1, int ksplit, int kskip)
{
  typedef ShapeMgO216w W;
  int gt = plane_preferred_gthreads(np0, np1, ksplit, kskip);
  if (const char* e = getenv("QB200_GROUP_THREADS")) { const int t = atoi(e); if (t >= 32 && t <= 256 && t % 32 == 0) gt = t; }
  // opt-in (QB200_PLANE_W=1): measured 14.8 ms against 14.6 ms for k_plane_s on the same 14 x 32 geometry -- the x phase is
  // too short for the warps to drift apart between the remaining barriers, so removing five of eight barriers bought nothing
  const char* ew = getenv("QB200_PLANE_W");
  const bool want_w = ew && ew[0] == '1';
  return (want_w && np0 == W::NP0 && np1 == W::NP1 && ksplit == W::YSPLIT && kskip == W::YSKIP && gt == W::GT) ? ShapeMgO216ww::PITCH : (np0 | 1);
}

// 0: generic kernel; > 0: index of the compiled shape that matches the plan (hmax = max |rod_h|)
int plane_select_static(const qb200_plan* p, int hmax)
{
  if (const char* e = getenv("QB200_NO_STATIC")) if (e[0] == '1') return 0;
  const DevPlan& d = p->d;
  typedef ShapeMgO216 S;
  if (d.np0 == S::NP0 && d.np1 == S::NP1 && d.ksplit == S::YSPLIT && d.kskip == S::YSKIP && hmax < S::XSPLIT &&
      p->plane_threads == S::NTHR && d.gthreads == S::GT && d.pitch0 == S::PITCH) return 1;
  typedef ShapeMgO216w W;
  if (d.np0 == W::NP0 && d.np1 == W::NP1 && d.ksplit == W::YSPLIT && d.kskip == W::YSKIP && hmax < W::XSPLIT &&
      p->plane_threads == W::NTHR && d.gthreads == W::GT && d.pitch0 == W::PITCH) return 2;
  typedef ShapeMgO216ww WW;
  if (d.np0 == WW::NP0 && d.np1 == WW::NP1 && d.ksplit == WW::YSPLIT && d.kskip == WW::YSKIP && hmax < WW::XSPLIT &&
      p->plane_threads == WW::NTHR && d.gthreads == WW::GT && d.pitch0 == WW::PITCH) return 3;
  return 0;
}

// ------------------------------------------------------------------------------------------------ tensor-memory kernel
// k_plane_t (plane_tmem.cuh) exists for the compiled MgO216 geometry; QB200_PLANE_T=0 keeps k_plane_s
typedef PlaneShape<112, 112, 26, 60, 26, 60, 14, 32, 113> ShapeMgO216t;
// warps (Y, X) per operation; QB200_T_HPSI / QB200_T_DENS = index into the lists below selects the alternative geometry.
// Measured (MgO216, profiles/r2t6_plane_t_warp_configs.txt): H psi 6.11 ms with (8, 8), 6.21 with (12, 8) at 96 registers;
// density 3.30 ms with (12, 4), 3.31 (8, 8), 3.34 (12, 8), 3.44 (12, 6): the kernel is bound by instruction issue (FP64 pipe
// 68 % + 21 % other instructions), not by the Y warps' critical path, so more warps do not help.  With the Good-Thomas passes
// (fewer Y instructions) the density runs 2.98 ms with (8, 8) against 3.14 with (12, 4): (8, 8) is the default for both
#define QB200_T_HPSI_LIST(F) F(0, 8, 8) F(1, 12, 8)
#define QB200_T_DENS_LIST(F) F(0, 8, 8) F(1, 12, 4)
static int t_cfg(const char* name) { const char* e = getenv(name); return e ? atoi(e) : 0; }

bool plane_t_wanted(const qb200_plan* p)
{
  if (const char* e = getenv("QB200_PLANE_T")) if (e[0] == '0') return false;
  typedef ShapeMgO216t T;
  const DevPlan& d = p->d;
  return p->fused && p->static_shape == 2 && d.np0 == T::NP0 && d.np1 == T::NP1 && d.ksplit == T::YSPLIT && d.kskip == T::YSKIP;
}
int plane_t_pitch() { return ShapeMgO216t::PITCH; }
void plane_t_xrange(int* xsplit, int* xskip) { *xsplit = ShapeMgO216t::XSPLIT; *xskip = ShapeMgO216t::XSKIP; }

int plane_t_setup(qb200_plan* p)
{
  typedef ShapeMgO216t T;
  p->plane_t = false;
  p->smem_plane_t = plane_t_smem<T>(p->d.nvec, p->d.ntzero);
  if (p->smem_plane_t + 64 > (size_t)p->max_smem) return QB200_OK;      // does not fit: k_plane_s stays
  // W_112^{b k1} for the thread-per-column passes (long double, as the packed tables)
  double tw[2 * 7 * 16];
  const long double twopi = 6.283185307179586476925286766559005768L;
  for (int b = 0; b < 7; b++)
    for (int k1 = 0; k1 < 16; k1++) {
      const int e = (b * k1) % T::NP1;
      tw[2 * (16 * b + k1)] = (double)cosl(twopi * e / T::NP1);
      tw[2 * (16 * b + k1) + 1] = (double)sinl(twopi * e / T::NP1);
    }
  QB_CUDA(cudaMemcpyToSymbol(c_ytw, tw, sizeof(tw)));
  int yrow[16 * 8] = { 0 };
  for (int k1 = 0; k1 < 16; k1++) for (int k2 = 0; k2 < 7; k2++) yrow[8 * k1 + k2] = ((49 * k1 + 64 * k2) % T::NP1) * T::NP0;
  QB_CUDA(cudaMemcpyToSymbol(c_yrow, yrow, sizeof(yrow)));
#define QB200_T_OPTIN_H(i, ny, nx) QB_CUDA(cudaFuncSetAttribute(k_plane_t<OP_HPSI, ShapeMgO216t, ny, nx>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p->smem_plane_t));
#define QB200_T_OPTIN_D(i, ny, nx) QB_CUDA(cudaFuncSetAttribute(k_plane_t<OP_DENSITY, ShapeMgO216t, ny, nx>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p->smem_plane_t));
  QB200_T_HPSI_LIST(QB200_T_OPTIN_H)
  QB200_T_DENS_LIST(QB200_T_OPTIN_D)
  // density with the CTA's rho plane in shared memory (k_plane_td): opt-in (QB200_T_DENS_SMEM=1).  Measured 3.53 ms against 3.29 ms
  // for the L2 reductions of k_plane_t<DENSITY> (MgO216): the reductions were not the limit, and single-buffering the kept rows
  // to make room for the plane costs more than they did
  p->plane_td = false;
  p->smem_plane_td = plane_td_smem<T>(p->d.nvec);
  {
    const char* e = getenv("QB200_T_DENS_SMEM");
    if (e && e[0] == '1' && p->smem_plane_td + 64 <= (size_t)p->max_smem) {
      QB_CUDA((cudaFuncSetAttribute(k_plane_td<ShapeMgO216t, 8, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p->smem_plane_td)));
      p->plane_td = true;
    }
  }
  p->plane_t = true;
  return QB200_OK;
}

template <class K> static int opt_in(K kernel, int bytes)
{
  QB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  return QB200_OK;
}

int plane_opt_in(qb200_plan* p)
{
  if ((int)p->smem_plane <= 48 * 1024) return QB200_OK;
  // per-kernel attribute shared by every plan of the process (wavefunction basis, density basis, ...): always the device
  // maximum, so that a later plan with a smaller plane cannot undercut an earlier one
  const int bytes = std::max((int)p->smem_plane, p->max_smem);
  if (p->static_shape == 1) {
    int rc;
    if ((rc = opt_in(k_plane_s<OP_HPSI, ShapeMgO216>, bytes)) || (rc = opt_in(k_plane_s<OP_DENSITY, ShapeMgO216>, bytes)) ||
        (rc = opt_in(k_plane_s<OP_BWD, ShapeMgO216>, bytes)) || (rc = opt_in(k_plane_s<OP_FWD, ShapeMgO216>, bytes))) return rc;
  }
  if (p->static_shape == 3) {
    int rc;
    if ((rc = opt_in(k_plane_w<OP_HPSI, ShapeMgO216ww>, bytes)) || (rc = opt_in(k_plane_w<OP_DENSITY, ShapeMgO216ww>, bytes)) ||
        (rc = opt_in(k_plane_w<OP_BWD, ShapeMgO216ww>, bytes)) || (rc = opt_in(k_plane_w<OP_FWD, ShapeMgO216ww>, bytes))) return rc;
  }
  if (p->static_shape == 2) {
    int rc;
    if ((rc = opt_in(k_plane_s<OP_HPSI, ShapeMgO216w>, bytes)) || (rc = opt_in(k_plane_s<OP_DENSITY, ShapeMgO216w>, bytes)) ||
        (rc = opt_in(k_plane_s<OP_BWD, ShapeMgO216w>, bytes)) || (rc = opt_in(k_plane_s<OP_FWD, ShapeMgO216w>, bytes))) return rc;
  }
  QB_CUDA(cudaFuncSetAttribute(k_plane2<OP_HPSI>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  QB_CUDA(cudaFuncSetAttribute(k_plane2<OP_DENSITY>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  QB_CUDA(cudaFuncSetAttribute(k_plane2<OP_BWD>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  QB_CUDA(cudaFuncSetAttribute(k_plane2<OP_FWD>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  return QB200_OK;
}

int launch_plane(qb200_plan* p, int op, dim3 grid, const double* v, double* f, const double* fac, int nunits, int zero_imag)
{
  const DevPlan& d = p->d;
  cplx* zt = (cplx*)p->zt;
  if (p->plane_t && (op == OP_DENSITY || (op == OP_HPSI && !zero_imag))) {
    if (op == OP_DENSITY && p->plane_td) {
      k_plane_td<ShapeMgO216t, 8, 8><<<grid, 512, p->smem_plane_td, p->stream>>>(d, zt, p->rho_part, fac, nunits);
      const cudaError_t e = cudaGetLastError();
      if (e != cudaSuccess) return cuda_fail(e, "k_plane_td launch", __FILE__, __LINE__);
      return QB200_OK;
    }
    const int cfg = t_cfg(op == OP_HPSI ? "QB200_T_HPSI" : "QB200_T_DENS");
#define QB200_T_LAUNCH_H(i, ny, nx) if (op == OP_HPSI && cfg == i) k_plane_t<OP_HPSI, ShapeMgO216t, ny, nx><<<grid, (ny + nx) * 32, p->smem_plane_t, p->stream>>>(d, zt, v, p->rho_part, fac, nunits, zero_imag);
#define QB200_T_LAUNCH_D(i, ny, nx) if (op == OP_DENSITY && cfg == i) k_plane_t<OP_DENSITY, ShapeMgO216t, ny, nx><<<grid, (ny + nx) * 32, p->smem_plane_t, p->stream>>>(d, zt, v, p->rho_part, fac, nunits, zero_imag);
    QB200_T_HPSI_LIST(QB200_T_LAUNCH_H)
    QB200_T_DENS_LIST(QB200_T_LAUNCH_D)
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "k_plane_t launch", __FILE__, __LINE__);
    return QB200_OK;
  }
  if (p->static_shape == 1) {
    typedef ShapeMgO216 S;
    switch (op) {
      case OP_HPSI: k_plane_s<OP_HPSI, S><<<grid, S::NTHR, p->smem_plane, p->stream>>>(d, zt, v, (cplx*)f, p->rho_part, fac, nunits, zero_imag); break;
      case OP_DENSITY: k_plane_s<OP_DENSITY, S><<<grid, S::NTHR, p->smem_plane, p->stream>>>(d, zt, v, (cplx*)f, p->rho_part, fac, nunits, zero_imag); break;
      case OP_BWD: k_plane_s<OP_BWD, S><<<grid, S::NTHR, p->smem_plane, p->stream>>>(d, zt, v, (cplx*)f, p->rho_part, fac, nunits, zero_imag); break;
      default: k_plane_s<OP_FWD, S><<<grid, S::NTHR, p->smem_plane, p->stream>>>(d, zt, v, (cplx*)f, p->rho_part, fac, nunits, zero_imag); break;
    }
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "k_plane_s launch", __FILE__, __LINE__);
    return QB200_OK;
  }
  if (p->static_shape == 3) {
    typedef ShapeMgO216ww S;
    switch (op) {
      case OP_HPSI: k_plane_w<OP_HPSI, S><<<grid, S::NTHR, p->smem_plane, p->stream>>>(d, zt, v, (cplx*)f, p->rho_part, fac, nunits, zero_imag); break;
      case OP_DENSITY: k_plane_w<OP_DENSITY, S><<<grid, S::NTHR, p->smem_plane, p->stream>>>(d, zt, v, (cplx*)f, p->rho_part, fac, nunits, zero_imag); break;
      case OP_BWD: k_plane_w<OP_BWD, S><<<grid, S::NTHR, p->smem_plane, p->stream>>>(d, zt, v, (cplx*)f, p->rho_part, fac, nunits, zero_imag); break;
      default: k_plane_w<OP_FWD, S><<<grid, S::NTHR, p->smem_plane, p->stream>>>(d, zt, v, (cplx*)f, p->rho_part, fac, nunits, zero_imag); break;
    }
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "k_plane_w launch", __FILE__, __LINE__);
    return QB200_OK;
  }
  if (p->static_shape == 2) {
    typedef ShapeMgO216w S;
    switch (op) {
      case OP_HPSI: k_plane_s<OP_HPSI, S><<<grid, S::NTHR, p->smem_plane, p->stream>>>(d, zt, v, (cplx*)f, p->rho_part, fac, nunits, zero_imag); break;
      case OP_DENSITY: k_plane_s<OP_DENSITY, S><<<grid, S::NTHR, p->smem_plane, p->stream>>>(d, zt, v, (cplx*)f, p->rho_part, fac, nunits, zero_imag); break;
      case OP_BWD: k_plane_s<OP_BWD, S><<<grid, S::NTHR, p->smem_plane, p->stream>>>(d, zt, v, (cplx*)f, p->rho_part, fac, nunits, zero_imag); break;
      default: k_plane_s<OP_FWD, S><<<grid, S::NTHR, p->smem_plane, p->stream>>>(d, zt, v, (cplx*)f, p->rho_part, fac, nunits, zero_imag); break;
    }
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "k_plane_s launch", __FILE__, __LINE__);
    return QB200_OK;
  }
  switch (op) {
    case OP_HPSI: k_plane2<OP_HPSI><<<grid, p->plane_threads, p->smem_plane, p->stream>>>(d, zt, v, (cplx*)f, p->rho_part, fac, nunits, zero_imag); break;
    case OP_DENSITY: k_plane2<OP_DENSITY><<<grid, p->plane_threads, p->smem_plane, p->stream>>>(d, zt, v, (cplx*)f, p->rho_part, fac, nunits, zero_imag); break;
    case OP_BWD: k_plane2<OP_BWD><<<grid, p->plane_threads, p->smem_plane, p->stream>>>(d, zt, v, (cplx*)f, p->rho_part, fac, nunits, zero_imag); break;
    default: k_plane2<OP_FWD><<<grid, p->plane_threads, p->smem_plane, p->stream>>>(d, zt, v, (cplx*)f, p->rho_part, fac, nunits, zero_imag); break;
  }
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "k_plane2 launch", __FILE__, __LINE__);
  return QB200_OK;
}

}  // namespace qb200
