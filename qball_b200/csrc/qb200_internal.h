// qball_b200/csrc/qb200_internal.h -- shared between the translation units of libqball_b200.so
#pragma once
#include <cuda_runtime.h>
#include <string>
#include <vector>
#include "../../include/qball_b200.h"
#include "fft_smem.cuh"

namespace qb200 {

void set_error(const std::string& s);
int cuda_fail(cudaError_t e, const char* what, const char* file, int line);
#define QB_CUDA(x) do { cudaError_t e__ = (x); if (e__ != cudaSuccess) return qb200::cuda_fail(e__, #x, __FILE__, __LINE__); } while (0)

bool is_device_ptr(const void* p);
// per-launch event timing (transform.cu)
void prof_begin(int cat, cudaStream_t s);
void prof_end(cudaStream_t s);
bool factorize(int n, FftDesc& d);
std::vector<double> packed_twiddle_table(const FftDesc& d);
std::vector<double> twiddle_table(int n);   // 2n doubles: cos, sin (2 pi j / n), long-double accurate

// device-side view of a plan (passed by value to kernels)
struct DevPlan {
  int np0, np1, np2;
  int nvec, nrods, ngw, is_real;
  int ntrans0, nkeep, ksplit, kskip;   // kept x-rows: [0,ksplit) and [ksplit+kskip, np1)
  int pitch0;                          // odd row pitch of a plane in shared memory
  int rb;                              // rods per z-column CTA
  int xb;                              // x columns per CTA in the split path
  FftDesc f0, f1, f2;
  const cplx *tw0, *tw1, *tw2;         // natural twiddle tables w_n^j (first-generation engine)
  const cplx *tw0p, *tw1p, *tw2p;      // packed per-pass twiddle tables (fft_desc.h)
  const int *rod_first, *rod_size, *rod_lmin;
  int gthreads;                        // threads per group in the plane kernel (fft_group.cuh)
  int nyrev_c;                         // 16-byte slots reserved for the yrev table in shared memory
  int ncolpos_c;                       // 16-byte slots for the colpos table in shared memory (0: read it from global)
  int stage_per;                       // plane kernel: column values per group staged ahead in dead rows (0: off)
  int exp;                             // QB200_EXP: timing experiments that BREAK results (bit 0 no rho reduction, 1 no v load, 2 no fill/scatter)
  const int *yrev;                     // natural y index of the first element of segment seg of the y mid pass
  const int *colpos;                   // per column iv: kp*pitch0 + (digit-reversed x position of hp)
  const int *colhk;                    // per column iv: hp + np0*kp
  const int *keepcols, *keeprowstart;  // split path: columns sorted by kept row; start offsets per kept row (nkeep+1)
  // second-generation z-column kernels (zcol_kernels.cuh); zb_*: backward, zf_*: forward
  const int *zq, *zqm;                 // per coefficient: digit-reversed z position | column << 12 (zqm: the -G image, real bases)
  int zb_rb, zb_cb, zb_cmax;           // rods per CTA, columns per tile, max coefficients per rod block
  int zf_rb, zf_cb, zf_cmax;
  // second-generation split xy stage (split_kernels.cuh)
  const int *xs_jr, *xs_x;             // per entry of keepcols: kept-row index, digit-reversed x position
  const int *yq;                       // natural y index held by position q after the y DIF transform
  // warp-owned plane kernel (k_plane_w): columns sorted by owner warp: staged address | plane position << 16, column index
  const int *wown, *wown_iv, *wown_start;
  // tensor-memory plane kernel (k_plane_t, plane_tmem.cuh): 16-bit positions in a buffer of kept rows only
  const unsigned short *tpos;          // per column iv: (kept-row index)*pitch + digit-reversed x position
  const unsigned short *tzero;         // positions inside the non-zero x range of a kept row that no column covers
  int ntzero;
  // density of Gamma-point REAL bases: a unit is a PAIR of states transformed as psi_1 + i psi_2 (SlaterDet.cc:858-899), whose
  // weights are fac[unit] (real part) and fac[unit + fac2off] (imaginary part); 0: one state per unit, one weight
  int fac2off;
};

enum { MODE_SINGLE = 0, MODE_PAIR = 1 };
#ifdef __CUDACC__
// weights of a density unit: only positive weights contribute (SlaterDet.cc:856, 905: states with zero occupation are skipped)
__device__ __forceinline__ bool fac_active(const DevPlan& P, const double* __restrict__ fac, int u)
{
  return fac[u] > 0.0 || (P.fac2off && fac[u + P.fac2off] > 0.0);
}
__device__ __forceinline__ double fac_first(const double* __restrict__ fac, int u) { const double f = fac[u]; return f > 0.0 ? f : 0.0; }
__device__ __forceinline__ double fac_second(const DevPlan& P, const double* __restrict__ fac, int u)
{
  const double f = fac[u + P.fac2off];          // fac2off == 0: the same weight for both parts (|psi|^2 of a complex state)
  return f > 0.0 ? f : 0.0;
}
#endif
enum { OP_HPSI = 0, OP_DENSITY = 1, OP_BWD = 2, OP_FWD = 3 };

}  // namespace qb200

struct qb200_plan;
namespace qb200 {
// plane.cu: the plane-fused xy stage (own translation unit: its register budget is 144, the other kernels' is 128)
int plane_opt_in(qb200_plan* p);
int plane_select_static(const qb200_plan* p, int hmax);
int plane_preferred_gthreads(int np0, int np1, int ksplit, int kskip);
int plane_preferred_pitch(int np0, int np1, int ksplit, int kskip);
int launch_plane(qb200_plan* p, int op, dim3 grid, const double* v, double* f, const double* fac, int nunits, int zero_imag);
bool plane_t_wanted(const qb200_plan* p);      // the compiled shape has a tensor-memory kernel and it is not switched off
int plane_t_pitch();
void plane_t_xrange(int* xsplit, int* xskip);
int plane_t_setup(qb200_plan* p);              // after d.tpos / d.tzero are uploaded: shared memory, constants, opt-in
// zcol_tmem.cu: z-column kernels with one thread per column in tensor memory (complex bases on the compiled 112-plane shape)
bool zcol_t_wanted(const qb200_plan* p, int lmax);
int zcol_t_setup(qb200_plan* p, const std::vector<int>& rod_first);
int launch_zbwd_t(qb200_plan* p, const double* c, size_t ldc, int nunits);
int launch_zfwd_t(qb200_plan* p, double* out, size_t ldc, int nunits, int accumulate, const double* kpg2, const double* cin, double scale);
// ycols_tmem.cu: y stage of the split xy path with the column in tensor memory (compiled 126 / 252 plane heights)
int ycols_t_setup(qb200_plan* p);
int launch_ycols_t(qb200_plan* p, int op, const double* v, const double* fac, int nunits, int zero_imag);
// ycols_tmem.cu: the whole xy stage of a 126 x 126 plane in one kernel (kept rows in shared memory, columns in tensor memory)
bool plane_f_wanted(const qb200_plan* p, int hmax);
int plane_f_pitch();
void plane_f_xrange(int* xsplit, int* xskip);
int plane_f_setup(qb200_plan* p);
int launch_plane_f(qb200_plan* p, int op, dim3 grid, const double* v, const double* fac, int nunits);
}

struct qb200_plan {
  int device;
  cudaStream_t stream;
  qb200::DevPlan d;
  std::vector<void*> owned;            // device allocations freed at destroy
  bool fused;                          // plane fits in shared memory
  size_t smem_z, smem_plane, smem_rows, smem_ycol;
  bool split2;                         // second-generation split xy kernels in use (planes larger than shared memory)
  int split_static;                    // 0 generic engine, 1 compiled 252 x 252 shape (gold benchmark)
  int xr_rowb, xr_smax;                // k_xrows2: rows per CTA, max column values per row block
  size_t smem_xr[2], smem_yc[4];       // dynamic shared memory of k_xrows2<+1/-1>, k_ycols2<OP>
  bool z2;                             // second-generation z-column kernels in use
  int z_static;                        // 0 run-time shape, 1 compiled 112-plane shape (MgO216)
  int z_threads;                       // threads per CTA of the v2 z kernels (256, or 128 with four CTAs per SM)
  size_t smem_zb[2], smem_zf[2];       // their dynamic shared memory, [MODE_SINGLE], [MODE_PAIR]
  int zslots_b[2], zslots_f[2];        // resident CTAs on the whole device
  long long ws_bytes;
  int batch;                           // units per batch
  double* zt;                          // [batch][np2][nvec] complex
  size_t zt_units;
  double* w;                           // split path: [batch][np2][nkeep][np0] complex
  size_t w_units;
  double* rho_part; size_t rho_part_elems;
  double* fac_dev; size_t fac_cap;
  std::vector<double> fac_host;        // host weights regrouped for pair units before their one upload
  // staging for host-pointer calls
  double *st_c, *st_cp, *st_v, *st_f, *st_kpg2; size_t st_c_cap, st_cp_cap, st_v_cap, st_f_cap, st_kpg2_cap;
  long long launches;
  int max_smem;
  int nsm;
  int plane_threads;
  int static_shape;                    // plane.cu: 0 generic kernel, > 0 compiled shape index
  bool plane_t;                        // H psi / density planes run k_plane_t (y direction in tensor memory)
  size_t smem_plane_t;
  bool plane_td;                       // density planes run k_plane_td (rho plane accumulated in shared memory)
  size_t smem_plane_td;
  bool zcol_t;                         // MODE_SINGLE z columns run k_zcol_bwd_t / k_zcol_fwd_t (zcol_tmem.cu)
  int zt_cmax, zt_nblk;                // coefficients of the longest 128-column block, number of blocks
  size_t smem_zt_b, smem_zt_f;
  bool plane_f;                        // planes too large for k_plane_t's two buffers but whose kept rows fit once: k_plane_f (si54p)
  size_t smem_plane_f;
  int ycols_t;                         // split path: 0 k_ycols2, 1 k_ycols_t on 126-row planes (si54p), 2 on 252-row planes (Au992)
  // pipelined host-pointer paths (hpsi.cu, qb200_compute_density): copy streams + events, and the identity of the host
  // coefficient block whose device copy sits in st_c (qb200_plan_set_coefficient_tag)
  cudaStream_t s_in, s_out;
  std::vector<cudaEvent_t> evs;
  const void* res_ptr; int res_ldc, res_nst; long long res_tag, next_tag;
  double *ex_a, *ex_b, *ex_c2; size_t ex_a_cap, ex_b_cap, ex_c2_cap;   // work blocks of qb200_exponential
  double* vh; size_t vh_cap;                                            // work of qb200_update_vhxc (vhxc.cuh)
};

namespace qb200 {
int plan_copy_streams(qb200_plan* p);                 // create s_in/s_out on first use
int plan_event(qb200_plan* p, size_t i, cudaEvent_t* ev);   // i-th event of the plan's pool (timing disabled)
bool plan_resident(const qb200_plan* p, const void* c, int ldc, int nst);
void plan_mark_resident(qb200_plan* p, const void* c, int ldc, int nst);
}
