// qball_b200/csrc/transform.cu -- plan construction and the FourierTransform / rs_mul_add / compute_density entry
// points of the C ABI (include/qball_b200.h).  Host code only builds tables and launches kernels; all arithmetic on
// wavefunction data happens in the kernels of transform_kernels.cuh.  There is no CPU fallback.
#include "transform_kernels.cuh"
#include "zcol_kernels.cuh"
#include "split_kernels.cuh"
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace qb200 {

static thread_local std::string g_err;
void set_error(const std::string& s) { g_err = s; }
int cuda_fail(cudaError_t e, const char* what, const char* file, int line)
{
  char buf[512];
  snprintf(buf, sizeof buf, "CUDA error %d (%s) in %s at %s:%d", (int)e, cudaGetErrorString(e), what, file, line);
  g_err = buf;
  if (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver || e == cudaErrorNoKernelImageForDevice) return QB200_ENODEV;
  if (e == cudaErrorMemoryAllocation) return QB200_ENOMEM;
  return QB200_ECUDA;
}

bool is_device_ptr(const void* p)
{
  if (!p) return false;
  cudaPointerAttributes a;
  cudaError_t e = cudaPointerGetAttributes(&a, p);
  if (e != cudaSuccess) { cudaGetLastError(); return false; }
  return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

bool factorize(int n, FftDesc& d)
{
  d = make_fft_desc(n);
  return d.nf > 0;
}

// packed per-pass twiddles (fft_desc.h): entry (t, k) of pass s = (cos, sin)(2 pi k t / len_s)
std::vector<double> packed_twiddle_table(const FftDesc& d)
{
  std::vector<double> t(2 * (size_t)d.twsize, 0.0);
  t[0] = 1.0;
  for (int s = 0; s < d.nf; s++) {
    const int r = d.r[s], len = d.len[s], m = len / r;
    if (m <= 1) continue;
    for (int tt = 0; tt < m; tt++)
      for (int k = 1; k < r; k++) {
        const long double a = 2.0L * 3.14159265358979323846264338327950288L * (long double)((long)k * tt % len) / (long double)len;
        const size_t i = (size_t)d.twoff[s] + (size_t)tt * (r - 1) + (k - 1);
        t[2 * i] = (double)cosl(a);
        t[2 * i + 1] = (double)sinl(a);
      }
  }
  return t;
}

std::vector<double> twiddle_table(int n)
{
  std::vector<double> t(2 * (size_t)n);
  for (int j = 0; j < n; j++) {
    const long double a = 2.0L * 3.14159265358979323846264338327950288L * (long double)j / (long double)n;
    t[2 * j] = (double)cosl(a);
    t[2 * j + 1] = (double)sinl(a);
  }
  return t;
}

struct ProfRec { int cat; cudaEvent_t a, b; bool open; };
static bool g_prof_on = false;
static std::vector<ProfRec> g_prof;
// several prof_begin calls before one launch put it into several categories; prof_end closes all of them
void prof_begin(int cat, cudaStream_t s)
{
  if (!g_prof_on) return;
  ProfRec r; r.cat = cat; r.open = true;
  cudaEventCreate(&r.a); cudaEventCreate(&r.b);
  cudaEventRecord(r.a, s);
  g_prof.push_back(r);
}
void prof_end(cudaStream_t s)
{
  if (!g_prof_on) return;
  for (size_t i = g_prof.size(); i > 0 && g_prof[i - 1].open; i--) {
    cudaEventRecord(g_prof[i - 1].b, s);
    g_prof[i - 1].open = false;
  }
}

template <class T> static int upload(qb200_plan* p, const std::vector<T>& h, const T** dptr)
{
  void* d = nullptr;
  size_t bytes = std::max<size_t>(h.size(), 1) * sizeof(T);
  QB_CUDA(cudaMalloc(&d, bytes));
  p->owned.push_back(d);
  if (!h.empty()) QB_CUDA(cudaMemcpy(d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
  *dptr = (const T*)d;
  return QB200_OK;
}

static int ensure(double** buf, size_t* cap, size_t elems)
{
  if (*cap >= elems && *buf) return QB200_OK;
  if (*buf) { cudaFree(*buf); *buf = nullptr; *cap = 0; }
  QB_CUDA(cudaMalloc((void**)buf, std::max<size_t>(elems, 1) * sizeof(double)));
  *cap = elems;
  return QB200_OK;
}

int plan_copy_streams(qb200_plan* p)
{
  if (!p->s_in) QB_CUDA(cudaStreamCreateWithFlags(&p->s_in, cudaStreamNonBlocking));
  if (!p->s_out) QB_CUDA(cudaStreamCreateWithFlags(&p->s_out, cudaStreamNonBlocking));
  return QB200_OK;
}
int plan_event(qb200_plan* p, size_t i, cudaEvent_t* ev)
{
  while (p->evs.size() <= i) {
    cudaEvent_t e;
    QB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    p->evs.push_back(e);
  }
  *ev = p->evs[i];
  return QB200_OK;
}
bool plan_resident(const qb200_plan* p, const void* c, int ldc, int nst)
{
  return p->next_tag != 0 && p->res_tag == p->next_tag && p->res_ptr == c && p->res_ldc == ldc && p->res_nst == nst && p->st_c;
}
void plan_mark_resident(qb200_plan* p, const void* c, int ldc, int nst)
{
  p->res_ptr = c; p->res_ldc = ldc; p->res_nst = nst; p->res_tag = p->next_tag;
}

// The attribute is per KERNEL, not per plan: two plans of one process (the wavefunction basis and the density basis of the
// same run) need different amounts for the same kernel, and the later, smaller request would undercut the earlier plan's
// launches ("invalid argument").  So every opt-in asks for the device's maximum; a launch still takes only what it needs.
template <class K> static int opt_in_smem(K kernel, size_t bytes)
{
  if (bytes <= 48 * 1024) return QB200_OK;
  int dev = 0, mx = 0;
  QB_CUDA(cudaGetDevice(&dev));
  QB_CUDA(cudaDeviceGetAttribute(&mx, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  QB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, std::max(mx, (int)bytes)));
  return QB200_OK;
}

}  // namespace qb200

using namespace qb200;

extern "C" const char* qb200_last_error(void) { return g_err.c_str(); }
extern "C" const char* qb200_version(void) { return "qball_b200 0.1 (sm_100a)"; }
extern "C" int qb200_device_count(void)
{
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

extern "C" int qb200_profile_enable(int on) { g_prof_on = on != 0; return QB200_OK; }
extern "C" int qb200_profile_read(double* ms, long long* count, int ncat)
{
  for (ProfRec& r : g_prof) {
    float t = 0.f;
    if (cudaEventSynchronize(r.b) == cudaSuccess && cudaEventElapsedTime(&t, r.a, r.b) == cudaSuccess && r.cat < ncat) {
      if (ms) ms[r.cat] += t;
      if (count) count[r.cat] += 1;
    }
    cudaEventDestroy(r.a); cudaEventDestroy(r.b);
  }
  g_prof.clear();
  cudaGetLastError();
  return QB200_OK;
}

static int configure_batch(qb200_plan* p)
{
  const DevPlan& d = p->d;
  const size_t zt_unit = (size_t)d.nvec * d.np2 * 16;
  const size_t w_unit = p->fused ? 0 : (size_t)d.nkeep * d.np0 * d.np2 * 16;
  long long b = p->ws_bytes / (long long)std::max<size_t>(zt_unit + w_unit, 1);
  if (b < 1) b = 1;
  if (b > 4096) b = 4096;
  p->batch = (int)b;
  return QB200_OK;
}

// fused path: how many of the `remaining` units to take in the next batch (<= p->batch) and over how many persistent
// CTAs per plane (grid.y = G <= maxG) to spread them, so that the np2*G CTAs fill whole waves of the nsm SMs (one CTA
// per SM) and every CTA gets the same number of units
static void plan_split(const qb200_plan* p, int remaining, int maxG, int* nb_out, int* G_out)
{
  const int np2 = p->d.np2, nsm = std::max(1, p->nsm);
  const int nbmax = std::min(p->batch, remaining);
  if (!p->fused && !p->plane_f) { *nb_out = nbmax; *G_out = 1; return; }
  double best = -1.0; int bnb = nbmax, bG = 1;
  if (const char* e = getenv("QB200_PLANE_G")) { const int G = atoi(e); if (G >= 1) { *nb_out = nbmax; *G_out = std::min(std::min(G, maxG), nbmax); return; } }
  const int nbmin = (nbmax == remaining) ? nbmax : std::max(1, (3 * nbmax) / 4);   // the last batch takes what is left
  for (int nb = nbmax; nb >= nbmin; nb--)
    for (int G = 1; G <= std::min(maxG, nb); G++) {
      const long rounds = ((long)np2 * G + nsm - 1) / nsm;
      const long upg = (nb + G - 1) / G;
      // a CTA pays ~0.4 unit-times of prologue (tables, first exposed fill) before its first unit
      // (the tensor-memory kernel is a two-stage pipeline over the units: one more unit-time to fill and drain)
      const double eff = ((double)nb * np2 / nsm) / ((double)rounds * (upg + (p->plane_t ? 1.2 : 0.4)));
      if (eff > best * (1.0 + 1e-9)) { best = eff; bnb = nb; bG = G; }
    }
  *nb_out = bnb; *G_out = bG;
}

static int ensure_work(qb200_plan* p, int units)
{
  const DevPlan& d = p->d;
  if ((size_t)units > p->zt_units) {
    if (p->zt) cudaFree(p->zt);
    p->zt = nullptr; p->zt_units = 0;
    QB_CUDA(cudaMalloc((void**)&p->zt, (size_t)units * d.nvec * d.np2 * 16 + 16));
    p->zt_units = units;
  }
  if (!p->fused && (size_t)units > p->w_units) {
    if (p->w) cudaFree(p->w);
    p->w = nullptr; p->w_units = 0;
    QB_CUDA(cudaMalloc((void**)&p->w, (size_t)units * d.nkeep * d.np0 * d.np2 * 16 + 16));
    p->w_units = units;
  }
  return QB200_OK;
}

typedef qb200::SplitShape<252, 252, 56, 140, 56, 140, 16, 8> ShapeAu992;
typedef qb200::SplitShape<126, 126, 29, 68, 29, 68, 16, 8> ShapeSi54p;    // examples/si54p at 65 Ry: 126^3 grid, |h|,|k| <= 28
typedef qb200::ZShape<112, 29, 26, 60> ZbMgO216;   // examples/MgO216: 112 planes, |l| <= 25; 29 / 22 columns per tile fill one wave
typedef qb200::ZShape<112, 22, 26, 60> ZfMgO216;
// the same shape in CTAs of 128 threads with tiles half as wide (four resident CTAs per SM instead of two; QB200_Z_THREADS=128)
typedef qb200::ZShape<112, 15, 26, 60> ZbMgO216n;
typedef qb200::ZShape<112, 11, 26, 60> ZfMgO216n;   // examples/gold_benchmark: 252 x 252 x 896 grid

template <class S> static bool split_shape_matches(const qb200::DevPlan& d, int hmax, int rowb)
{
  return d.np0 == S::NP0 && d.np1 == S::NP1 && d.ksplit == S::YSPLIT && d.kskip == S::YSKIP && hmax < S::XSPLIT && d.xb == S::XB &&
         rowb == S::ROWB;
}

template <class S> static int split_opt_in(qb200_plan* p)
{
  int rc;
  if ((rc = opt_in_smem(k_xrows2<+1, S>, p->smem_xr[0])) || (rc = opt_in_smem(k_xrows2<-1, S>, p->smem_xr[1])) ||
      (rc = opt_in_smem(k_ycols2<OP_HPSI, S>, p->smem_yc[OP_HPSI])) || (rc = opt_in_smem(k_ycols2<OP_DENSITY, S>, p->smem_yc[OP_DENSITY])) ||
      (rc = opt_in_smem(k_ycols2<OP_BWD, S>, p->smem_yc[OP_BWD])) || (rc = opt_in_smem(k_ycols2<OP_FWD, S>, p->smem_yc[OP_FWD]))) return rc;
  return QB200_OK;
}

// ------------------------------------------------------------------------------------------------ z-column kernels v2
// Picks, for each direction, the number of columns per tile (cb; rb = rods per CTA) so that the persistent grid
// (rod blocks x G) fills the resident-CTA slots of the device as evenly as possible, and records shared-memory sizes.
namespace {
struct Z2Choice { int cb, rb, cmax, ctas; size_t smem[2]; double cost; };

int z2_cmax(const std::vector<int>& first, int ngw, int nrods, int rb)
{
  int cmax = 1;
  for (int r0 = 0; r0 < nrods; r0 += rb) {
    const int r1 = std::min(r0 + rb, nrods);
    cmax = std::max(cmax, (r1 < nrods ? first[r1] : ngw) - first[r0]);
  }
  return (cmax + 3) & ~3;
}

bool z2_pick(const qb200_plan* p, const std::vector<int>& first, bool fwd, size_t smem_sm, int force_cb, Z2Choice* out)
{
  const DevPlan& d = p->d;
  const int per = d.is_real ? 2 : 1;
  bool found = false;
  Z2Choice best = {};
  // tiles of at least 8 columns (128-byte runs of a zt row); very long columns (the 896 planes of the gold benchmark: 14 KB per
  // column) only fit 3-4 per tile next to the staging buffers -- still faster than the first-generation kernels (Au992, 64
  // states: z stages 33.2 -> 25.9 ms).  QB200_Z2_MINCB overrides.
  int cbmin = d.np2 > 512 ? std::max(per, 3) : 8;
  if (const char* e = getenv("QB200_Z2_MINCB")) cbmin = std::max(per, atoi(e));
  for (int cb = 32; cb >= cbmin; cb--) {
    if (cb % per) continue;
    if (force_cb > 0 && cb != force_cb) continue;
    Z2Choice c;
    c.cb = cb; c.rb = cb / per;
    c.cmax = z2_cmax(first, d.ngw, d.nrods, c.rb);
    const size_t pitch = (size_t)(cb | 1), tile = (size_t)d.np2 * pitch * 16, tw = (size_t)d.f2.twsize * 16;
    for (int mode = 0; mode < 2; mode++) {
      const size_t cper = mode == MODE_PAIR ? 2 : 1;
      c.smem[mode] = fwd ? tw + 2 * tile + (size_t)c.cmax * (8 + 4 + 4)
                         : tw + tile + 2 * cper * (size_t)c.cmax * 16 + (size_t)c.cmax * 8;
    }
    const size_t need = c.smem[d.is_real ? MODE_PAIR : MODE_SINGLE];
    if (need > (size_t)p->max_smem) continue;
    c.ctas = (int)std::min<size_t>(p->z_threads == 128 ? 4 : 2, smem_sm / (need + 1024));
    if (c.ctas < 1) continue;
    const long slots = (long)p->nsm * c.ctas;
    const long nzb = (d.nrods + c.rb - 1) / c.rb;
    const long G = std::max(1l, slots / nzb);
    const long waves = (nzb * G + slots - 1) / slots;
    c.cost = (double)c.rb * c.ctas * waves / (double)G * (c.ctas == 1 ? 1.15 : 1.0) * (cb < 16 ? 1.25 : 1.0);
    if (!found || c.cost < best.cost * (1.0 - 1e-9)) { best = c; found = true; }
  }
  if (found) *out = best;
  return found;
}
}  // namespace

static int configure_z2(qb200_plan* p, const std::vector<int>& first, size_t smem_sm, int lmax)
{
  DevPlan& d = p->d;
  p->z2 = false;
  if (const char* e = getenv("QB200_Z2")) if (e[0] == '0') return QB200_OK;
  // zq packs (digit-reversed z) | (column << 12) in one int: 12 bits of z, 19 bits of column
  if (d.np2 > 4095 || d.nvec >= (1 << 19)) return QB200_OK;
  int fb = 0, ff = 0;
  p->z_threads = 256;
  if (const char* e = getenv("QB200_Z_THREADS")) if (atoi(e) == 128) p->z_threads = 128;
  if (const char* e = getenv("QB200_ZB_COLS")) fb = atoi(e);
  if (const char* e = getenv("QB200_ZF_COLS")) ff = atoi(e);
  // compiled shape (MgO216's 112 planes): its tile widths are fixed at compile time
  p->z_static = 0;
  {
    const char* ns = getenv("QB200_NO_STATIC");
    if (!(ns && ns[0] == '1') && !d.is_real && d.np2 == ZbMgO216::NP2 && lmax < ZbMgO216::ZSPLIT && fb == 0 && ff == 0) {
      p->z_static = 1; fb = ZbMgO216::CB; ff = ZfMgO216::CB;
      if (p->z_threads == 128) { fb = ZbMgO216n::CB; ff = ZfMgO216n::CB; }
    }
  }
  Z2Choice b, f;
  if (!z2_pick(p, first, false, smem_sm, fb, &b) || !z2_pick(p, first, true, smem_sm, ff, &f)) { p->z_static = 0; return QB200_OK; }
  d.zb_cb = b.cb; d.zb_rb = b.rb; d.zb_cmax = b.cmax;
  d.zf_cb = f.cb; d.zf_rb = f.rb; d.zf_cmax = f.cmax;
  int rc;
  for (int mode = 0; mode < 2; mode++) { p->smem_zb[mode] = b.smem[mode]; p->smem_zf[mode] = f.smem[mode]; }
  if ((rc = opt_in_smem(k_zcol_bwd2<MODE_SINGLE, DynZ>, p->smem_zb[0])) || (rc = opt_in_smem(k_zcol_bwd2<MODE_PAIR, DynZ>, p->smem_zb[1])) ||
      (rc = opt_in_smem(k_zcol_fwd2<MODE_SINGLE, DynZ>, p->smem_zf[0])) || (rc = opt_in_smem(k_zcol_fwd2<MODE_PAIR, DynZ>, p->smem_zf[1]))) return rc;
  if (p->z_static == 1 &&
      ((rc = opt_in_smem(k_zcol_bwd2<MODE_SINGLE, ZbMgO216>, p->smem_zb[0])) || (rc = opt_in_smem(k_zcol_fwd2<MODE_SINGLE, ZfMgO216>, p->smem_zf[0])) ||
       (rc = opt_in_smem(k_zcol_bwd2<MODE_SINGLE, ZbMgO216n>, p->smem_zb[0])) || (rc = opt_in_smem(k_zcol_fwd2<MODE_SINGLE, ZfMgO216n>, p->smem_zf[0])))) return rc;
  int n = 0;
  QB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_zcol_bwd2<MODE_SINGLE, DynZ>, p->z_threads, p->smem_zb[0])); p->zslots_b[0] = std::max(1, n) * p->nsm;
  QB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_zcol_bwd2<MODE_PAIR, DynZ>, p->z_threads, p->smem_zb[1])); p->zslots_b[1] = std::max(1, n) * p->nsm;
  QB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_zcol_fwd2<MODE_SINGLE, DynZ>, p->z_threads, p->smem_zf[0])); p->zslots_f[0] = std::max(1, n) * p->nsm;
  QB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_zcol_fwd2<MODE_PAIR, DynZ>, p->z_threads, p->smem_zf[1])); p->zslots_f[1] = std::max(1, n) * p->nsm;
  p->z2 = true;
  return QB200_OK;
}

extern "C" int qb200_plan_create(qb200_plan** out, int device, int np0, int np1, int np2, int nrods, const int* rod_h,
                                 const int* rod_k, const int* rod_lmin, const int* rod_size, int is_real, int idxmin1,
                                 int idxmax1)
{
  if (!out || np0 < 1 || np1 < 1 || np2 < 1 || nrods < 1 || !rod_h || !rod_k || !rod_lmin || !rod_size) {
    set_error("qb200_plan_create: bad argument");
    return QB200_EINVAL;
  }
  *out = nullptr;
  int ndev = 0;
  QB_CUDA(cudaGetDeviceCount(&ndev));
  if (device < 0 || device >= ndev) { set_error("qb200_plan_create: no such CUDA device"); return QB200_ENODEV; }
  QB_CUDA(cudaSetDevice(device));
  qb200_plan* p = new qb200_plan();
  memset(&p->d, 0, sizeof(p->d));
  p->device = device; p->stream = 0; p->zt = nullptr; p->zt_units = 0; p->w = nullptr; p->w_units = 0;
  p->rho_part = nullptr; p->rho_part_elems = 0; p->fac_dev = nullptr; p->fac_cap = 0;
  p->st_c = p->st_cp = p->st_v = p->st_f = p->st_kpg2 = nullptr;
  p->st_c_cap = p->st_cp_cap = p->st_v_cap = p->st_f_cap = p->st_kpg2_cap = 0;
  p->launches = 0;
  p->ex_a = p->ex_b = p->ex_c2 = nullptr; p->ex_a_cap = p->ex_b_cap = p->ex_c2_cap = 0;
  p->vh = nullptr; p->vh_cap = 0;
  p->s_in = p->s_out = nullptr; p->res_ptr = nullptr; p->res_ldc = p->res_nst = 0; p->res_tag = p->next_tag = 0;
  DevPlan& d = p->d;
  d.np0 = np0; d.np1 = np1; d.np2 = np2; d.nrods = nrods; d.is_real = is_real ? 1 : 0;
  d.nvec = is_real ? 2 * nrods - 1 : nrods;                       // FourierTransform.cc:186-197
  d.ntrans0 = std::max(std::abs(idxmax1), std::abs(idxmin1)) + 1;  // FourierTransform.cc:202
  if (2 * d.ntrans0 >= np1) { d.nkeep = np1; d.ksplit = np1; d.kskip = 0; }
  else { d.nkeep = 2 * d.ntrans0; d.ksplit = d.ntrans0; d.kskip = np1 - 2 * d.ntrans0; }
  d.pitch0 = plane_preferred_pitch(np0, np1, d.ksplit, d.kskip);
  if (!factorize(np0, d.f0) || !factorize(np1, d.f1) || !factorize(np2, d.f2)) {
    set_error("qb200_plan_create: grid length not of the form 2^a 3^b 5^c 7^d 11^e"); delete p; return QB200_EUNSUPPORTED;
  }
  for (const FftDesc* f : { &d.f0, &d.f1, &d.f2 })
    if (f->n / f->r[f->nf - 1] > 256) { set_error("qb200_plan_create: grid length too large"); delete p; return QB200_EUNSUPPORTED; }
  if (is_real && (rod_h[0] != 0 || rod_k[0] != 0 || rod_lmin[0] != 0)) {
    set_error("qb200_plan_create: real basis requires rod(0,0) first with lmin 0 (Basis.cc:637-651)"); delete p; return QB200_EINVAL;
  }
  // tables
  std::vector<int> first(nrods), size(rod_size, rod_size + nrods), lmin(rod_lmin, rod_lmin + nrods);
  int ngw = 0;
  for (int r = 0; r < nrods; r++) {
    first[r] = ngw; ngw += rod_size[r];
    const int lo = rod_lmin[r], hi = rod_lmin[r] + rod_size[r] - 1;
    if (rod_size[r] < 1 || rod_size[r] > np2 || lo <= -np2 || hi >= np2 || std::abs(rod_h[r]) >= np0 || std::abs(rod_k[r]) >= np1) {
      set_error("qb200_plan_create: rod does not fit the grid"); delete p; return QB200_EINVAL;
    }
  }
  d.ngw = ngw;
  std::vector<int> colpos(d.nvec), colhk(d.nvec), xpos(np0), yrev;
  for (int q = 0; q < np0; q++) xpos[digit_reverse(d.f0, q)] = q;   // x-rows are transformed DIT: scatter to digit-reversed x
  {
    const int rl = d.f1.r[d.f1.nf - 1];
    for (int seg = 0; seg < np1 / rl; seg++) yrev.push_back(digit_reverse(d.f1, seg * rl));
    d.nyrev_c = ((int)yrev.size() * 4 + 15) / 16;
  }
  auto put = [&](int iv, int hp, int kp) { colpos[iv] = kp * d.pitch0 + xpos[hp]; colhk[iv] = hp + np0 * kp; };
  for (int r = 0; r < nrods; r++) {                               // FourierTransform.cc:361-430, 484-506
    int hp = rod_h[r], kp = rod_k[r];
    if (hp < 0) hp += np0;
    if (kp < 0) kp += np1;
    if (!is_real) { put(r, hp, kp); continue; }
    if (r == 0) { put(0, 0, 0); continue; }
    int hm = -hp, km = -kp;
    if (hm < 0) hm += np0;
    if (km < 0) km += np1;
    put(2 * r - 1, hp, kp);
    put(2 * r, hm, km);
  }
  // every column must sit on a kept row (true by construction of ntrans0)
  std::vector<int> keepcols(d.nvec), rowstart(d.nkeep + 1, 0);
  {
    std::vector<std::vector<int> > by(d.nkeep);
    for (int iv = 0; iv < d.nvec; iv++) {
      const int kp = colhk[iv] / np0;
      int jr;
      if (kp < d.ksplit) jr = kp;
      else if (kp >= d.ksplit + d.kskip) jr = kp - d.kskip;
      else { set_error("qb200_plan_create: rod outside the kept rows (idxmin1/idxmax1 inconsistent)"); delete p; return QB200_EINVAL; }
      by[jr].push_back(iv);
    }
    int k = 0;
    for (int jr = 0; jr < d.nkeep; jr++) { rowstart[jr] = k; for (int iv : by[jr]) keepcols[k++] = iv; }
    rowstart[d.nkeep] = k;
  }
  std::vector<int> xs_jr(d.nvec), xs_x(d.nvec), yq(np1);
  for (int jr = 0; jr < d.nkeep; jr++)
    for (int i = rowstart[jr]; i < rowstart[jr + 1]; i++) { xs_jr[i] = jr; xs_x[i] = xpos[colhk[keepcols[i]] % np0]; }
  for (int q = 0; q < np1; q++) yq[q] = digit_reverse(d.f1, q);
  int rc;
  if ((rc = upload(p, xs_jr, &d.xs_jr)) || (rc = upload(p, xs_x, &d.xs_x)) || (rc = upload(p, yq, &d.yq))) { qb200_plan_destroy(p); return rc; }
  if ((rc = upload(p, first, &d.rod_first)) || (rc = upload(p, size, &d.rod_size)) || (rc = upload(p, lmin, &d.rod_lmin)) ||
      (rc = upload(p, colpos, &d.colpos)) || (rc = upload(p, yrev, &d.yrev)) || (rc = upload(p, colhk, &d.colhk)) || (rc = upload(p, keepcols, &d.keepcols)) ||
      (rc = upload(p, rowstart, &d.keeprowstart))) { qb200_plan_destroy(p); return rc; }
  {
    const double* t;
    if ((rc = upload(p, twiddle_table(np0), &t))) { qb200_plan_destroy(p); return rc; }
    d.tw0 = (const cplx*)t;
    if ((rc = upload(p, twiddle_table(np1), &t))) { qb200_plan_destroy(p); return rc; }
    d.tw1 = (const cplx*)t;
    if ((rc = upload(p, twiddle_table(np2), &t))) { qb200_plan_destroy(p); return rc; }
    d.tw2 = (const cplx*)t;
    if ((rc = upload(p, packed_twiddle_table(d.f0), &t))) { qb200_plan_destroy(p); return rc; }
    d.tw0p = (const cplx*)t;
    if ((rc = upload(p, packed_twiddle_table(d.f1), &t))) { qb200_plan_destroy(p); return rc; }
    d.tw1p = (const cplx*)t;
    if ((rc = upload(p, packed_twiddle_table(d.f2), &t))) { qb200_plan_destroy(p); return rc; }
    d.tw2p = (const cplx*)t;
  }
  // launch geometry
  cudaDeviceProp prop;
  { const cudaError_t e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) { const int rcp = cuda_fail(e, "cudaGetDeviceProperties", __FILE__, __LINE__); qb200_plan_destroy(p); return rcp; } }
  p->max_smem = (int)prop.sharedMemPerBlockOptin;
  p->nsm = prop.multiProcessorCount;
  {
    // second-generation z-column kernels: shared-memory position of every coefficient (digit-reversed z | column << 12)
    std::vector<int> zpos(np2), zq(ngw), zqm(is_real ? ngw : 0);
    for (int q = 0; q < np2; q++) zpos[digit_reverse(d.f2, q)] = q;
    for (int r = 0; r < nrods; r++) {
      const int col = is_real ? (r == 0 ? 0 : 2 * r - 1) : r, colm = r == 0 ? col : col + 1;
      for (int i = 0; i < size[r]; i++) {
        const int l = lmin[r] + i;
        const int izp = l < 0 ? l + np2 : l, izm = l > 0 ? np2 - l : -l;
        zq[first[r] + i] = zpos[izp] | (col << 12);
        if (is_real) zqm[first[r] + i] = zpos[izm] | (colm << 12);
      }
    }
    if ((rc = upload(p, zq, &d.zq)) || (rc = upload(p, zqm, &d.zqm))) { qb200_plan_destroy(p); return rc; }
    int lmax = 0;
    for (int r = 0; r < nrods; r++) lmax = std::max(lmax, std::max(std::abs(lmin[r]), std::abs(lmin[r] + size[r] - 1)));
    if ((rc = configure_z2(p, first, prop.sharedMemPerMultiprocessor, lmax))) { qb200_plan_destroy(p); return rc; }
    p->zcol_t = false;
    if (p->z2 && zcol_t_wanted(p, lmax) && (rc = zcol_t_setup(p, first))) { qb200_plan_destroy(p); return rc; }
  }
  const size_t pitch2 = (size_t)(np2 | 1);
  // columns per CTA of the first-generation z kernels: two CTAs per SM (2 x 104 KB), and an EVEN number of columns so that
  // the runs of a zt row a CTA writes / reads are whole 32-byte sectors (long columns: 6 x 16 B = 96 B for the 896 planes of
  // the gold benchmark instead of 5 x 16 B = 80 B straddling sectors; QB200_Z1_COLS overrides)
  int ncolmax = (int)std::min<size_t>(32, (104 * 1024) / (pitch2 * 16));
  if (ncolmax > 2) ncolmax &= ~1;
  if (const char* e = getenv("QB200_Z1_COLS")) { const int v = atoi(e); if (v >= 1 && v <= 32) ncolmax = v; }
  if (ncolmax < (is_real ? 2 : 1)) ncolmax = is_real ? 2 : 1;
  d.rb = is_real ? std::max(1, ncolmax / 2) : ncolmax;
  const int zcols = is_real ? 2 * d.rb : d.rb;
  p->smem_z = ((size_t)np2 + (size_t)zcols * pitch2) * 16;
  d.ncolpos_c = (d.nvec * 4 + 15) / 16;
  p->smem_plane = ((size_t)d.f0.twsize + d.f1.twsize + d.nyrev_c + d.ncolpos_c + (size_t)np1 * d.pitch0) * 16;
  if (p->smem_plane > (size_t)p->max_smem) {      // colpos stays in global memory
    d.ncolpos_c = 0;
    p->smem_plane = ((size_t)d.f0.twsize + d.f1.twsize + d.nyrev_c + (size_t)np1 * d.pitch0) * 16;
  }
  const char* force_split = getenv("QB200_FORCE_SPLIT");
  p->fused = p->smem_plane <= (size_t)p->max_smem && !(force_split && force_split[0] == '1');
  d.xb = 16;
  while (d.xb > 1 && ((size_t)np1 + (size_t)np1 * d.xb) * 16 > 96 * 1024) d.xb /= 2;
  p->smem_ycol = ((size_t)np1 + (size_t)np1 * d.xb) * 16;
  int rowb = (int)std::max<size_t>(1, (64 * 1024) / ((size_t)d.pitch0 * 16));
  rowb = std::min(rowb, d.nkeep);
  p->smem_rows = ((size_t)np0 + (size_t)rowb * d.pitch0) * 16;
  if (p->smem_z > (size_t)p->max_smem || p->smem_ycol > (size_t)p->max_smem || p->smem_rows > (size_t)p->max_smem) {
    set_error("qb200_plan_create: grid too large for shared-memory staging"); qb200_plan_destroy(p); return QB200_EUNSUPPORTED;
  }
  if ((rc = opt_in_smem(k_zcol_bwd<MODE_SINGLE>, p->smem_z)) || (rc = opt_in_smem(k_zcol_bwd<MODE_PAIR>, p->smem_z)) ||
      (rc = opt_in_smem(k_zcol_fwd<MODE_SINGLE>, p->smem_z)) || (rc = opt_in_smem(k_zcol_fwd<MODE_PAIR>, p->smem_z))) {
    qb200_plan_destroy(p); return rc;
  }
  p->split2 = false;
  if (!p->fused) {
    // second-generation split kernels: rows per x CTA sized for two resident CTAs, shared-memory budgets per kernel
    const size_t rowbytes = (size_t)d.pitch0 * 16;
    int rowb = (int)std::min<size_t>(d.nkeep, std::max<size_t>(1, (40 * 1024) / rowbytes));
    if (rowb > 8) rowb = 8;
    // (12 rows per CTA for the 252-wide rows of the gold benchmark, compiled in: xy 73.8 ms against 72.5 ms per 64 states -- no gain)
    if (const char* e = getenv("QB200_XR_ROWB")) { const int r = atoi(e); if (r >= 1 && r <= d.nkeep) rowb = r; }
    int smax = 4;
    for (int jr0 = 0; jr0 < d.nkeep; jr0 += rowb) smax = std::max(smax, rowstart[std::min(jr0 + rowb, d.nkeep)] - rowstart[jr0]);
    smax = (smax + 3) & ~3;
    p->xr_rowb = rowb; p->xr_smax = smax;
    p->smem_xr[0] = (size_t)d.f0.twsize * 16 + (size_t)rowb * rowbytes + (size_t)smax * (16 + 8);
    p->smem_xr[1] = (size_t)d.f0.twsize * 16 + 2 * (size_t)rowb * rowbytes + (size_t)smax * 8;
    const size_t ybase = (size_t)d.f1.twsize * 16 + (size_t)((np1 + 3) & ~3) * 4 + (size_t)np1 * (d.xb | 1) * 16 + (size_t)d.nkeep * d.xb * 16;
    for (int op = 0; op < 4; op++) p->smem_yc[op] = ybase + (op == OP_DENSITY ? (size_t)np1 * d.xb * 8 : 0);
    const char* e2 = getenv("QB200_SPLIT2");
    size_t need = std::max(p->smem_xr[0], p->smem_xr[1]);
    for (int op = 0; op < 4; op++) need = std::max(need, p->smem_yc[op]);
    p->split2 = !(e2 && e2[0] == '0') && need <= (size_t)p->max_smem;
    if (p->split2 &&
        ((rc = opt_in_smem(k_xrows2<+1, DynSplit>, p->smem_xr[0])) || (rc = opt_in_smem(k_xrows2<-1, DynSplit>, p->smem_xr[1])) ||
         (rc = opt_in_smem(k_ycols2<OP_HPSI, DynSplit>, p->smem_yc[OP_HPSI])) || (rc = opt_in_smem(k_ycols2<OP_DENSITY, DynSplit>, p->smem_yc[OP_DENSITY])) ||
         (rc = opt_in_smem(k_ycols2<OP_BWD, DynSplit>, p->smem_yc[OP_BWD])) || (rc = opt_in_smem(k_ycols2<OP_FWD, DynSplit>, p->smem_yc[OP_FWD])))) {
      qb200_plan_destroy(p); return rc;
    }
    // compiled shapes: the gold benchmark's 252 x 252 planes and si54p's 126 x 126 planes
    p->split_static = 0;
    {
      int hmax = 0;
      for (int r = 0; r < nrods; r++) hmax = std::max(hmax, std::abs(rod_h[r]));
      const char* ns = getenv("QB200_NO_STATIC");
      if (p->split2 && !(ns && ns[0] == '1')) {
        if (split_shape_matches<ShapeAu992>(d, hmax, rowb)) { p->split_static = 1; rc = split_opt_in<ShapeAu992>(p); }
        else if (split_shape_matches<ShapeSi54p>(d, hmax, rowb)) { p->split_static = 2; rc = split_opt_in<ShapeSi54p>(p); }
        if (rc) { qb200_plan_destroy(p); return rc; }
      }
    }
  }
  if (p->fused) {
    // (opt-in for the plane kernels happens after the thread geometry is chosen, below)
  } else {
    if ((rc = opt_in_smem(k_xrows<+1>, p->smem_rows)) || (rc = opt_in_smem(k_xrows<-1>, p->smem_rows)) ||
        (rc = opt_in_smem(k_ycols<OP_HPSI>, p->smem_ycol)) || (rc = opt_in_smem(k_ycols<OP_DENSITY>, p->smem_ycol)) ||
        (rc = opt_in_smem(k_ycols<OP_BWD>, p->smem_ycol)) || (rc = opt_in_smem(k_ycols<OP_FWD>, p->smem_ycol))) {
      qb200_plan_destroy(p); return rc;
    }
  }
  p->ycols_t = 0;
  if ((rc = ycols_t_setup(p))) { qb200_plan_destroy(p); return rc; }
  // plane kernel geometry: groups of gthreads threads own blocks of 8 rows / 8 columns (fft_group.cuh); pick the number of
  // groups (<= 7 x 64 = 448 threads so that the radix-16 pass keeps its ~112 registers) that needs the fewest rounds
  {
    d.gthreads = plane_preferred_gthreads(np0, np1, d.ksplit, d.kskip);
    if (const char* e = getenv("QB200_GROUP_THREADS")) { const int t = atoi(e); if (t >= 32 && t <= 256 && t % 32 == 0) d.gthreads = t; }
    const int maxg = std::max(1, std::min(15, 448 / d.gthreads));
    const int by = (np0 + 7) / 8, bx = (d.nkeep + 7) / 8 + ((d.nkeep < np1 && d.ksplit % 8) ? 1 : 0);
    const double wy = (double)np1, wx = (double)np0;
    auto cost = [&](int g) { return ((by + g - 1) / g) * wy * 8 + ((bx + g - 1) / g) * wx * 8; };
    int best = maxg;
    for (int g = maxg; g >= 1; g--) if (cost(g) < cost(best) * (1.0 - 1e-9)) best = g;
    p->plane_threads = best * d.gthreads;
    if (const char* e = getenv("QB200_PLANE_THREADS")) {
      const int t = atoi(e);
      if (t >= d.gthreads && t <= 448 && t % d.gthreads == 0 && t / d.gthreads <= 15) p->plane_threads = t;
    }
    // staging of the next unit's column values in the dead (not kept) rows of each group's first column block
    const int ngrp = p->plane_threads / d.gthreads;
    const int per = (((d.nvec + ngrp - 1) / ngrp) + 7) / 8 * 8;
    d.stage_per = (per <= d.kskip * 8 && ngrp * 8 <= np0) ? per : 0;
    if (const char* e = getenv("QB200_NO_STAGE")) if (e[0] == '1') d.stage_per = 0;
    d.exp = 0;
    if (const char* e = getenv("QB200_EXP")) d.exp = atoi(e);      // timing experiments only: results are wrong when set
  }
  // 16-bit positions in a buffer that holds the kept rows only (pitch tp): per column its (kept row, digit-reversed x), and the
  // positions inside the non-zero x range [0,xs) + [xs+xk,np0) of a kept row that no column covers (they must read as zero)
  auto build_kept_row_tables = [&](int tp, int xs, int xk) -> int {
    std::vector<unsigned short> tpos(d.nvec), tzero;
    std::vector<char> covered((size_t)d.nkeep * tp, 0);
    for (int iv = 0; iv < d.nvec; iv++) {
      const int hp = colhk[iv] % np0, kp = colhk[iv] / np0, jr = kp < d.ksplit ? kp : kp - d.kskip;
      tpos[iv] = (unsigned short)(jr * tp + xpos[hp]);
      covered[tpos[iv]] = 1;
    }
    for (int jr = 0; jr < d.nkeep; jr++)
      for (int hp = 0; hp < np0; hp++)
        if ((hp < xs || hp >= xs + xk) && !covered[(size_t)jr * tp + xpos[hp]]) tzero.push_back((unsigned short)(jr * tp + xpos[hp]));
    d.ntzero = (int)tzero.size();
    int r;
    if ((r = upload(p, tpos, &d.tpos)) || (r = upload(p, tzero, &d.tzero))) return r;
    return QB200_OK;
  };
  p->static_shape = 0;
  if (p->fused) {
    int hmax = 0;
    for (int r = 0; r < nrods; r++) hmax = std::max(hmax, std::abs(rod_h[r]));
    p->static_shape = plane_select_static(p, hmax);
    // warp-owned variant of the compiled MgO216 geometry: tables sorted by the warp that owns a column's kept row
    if (p->static_shape == 3 && !(p->smem_plane + (size_t)d.ncolpos_c * 16 <= (size_t)p->max_smem && d.ncolpos_c > 0 && d.stage_per >= 0)) p->static_shape = 0;
    if (p->static_shape == 3 && p->smem_plane + (size_t)d.ncolpos_c * 16 <= (size_t)p->max_smem && d.ncolpos_c > 0) {
      const int ngrp = p->plane_threads / d.gthreads, half = ngrp / 2, rbx = (d.ksplit + half - 1) / half, per = d.stage_per;
      std::vector<std::vector<int> > by(ngrp);
      for (int iv = 0; iv < d.nvec; iv++) {
        const int kp = colhk[iv] / np0;
        by[kp < d.ksplit ? kp / rbx : half + (kp - d.ksplit - d.kskip) / rbx].push_back(iv);
      }
      std::vector<int> wown, wiv, wstart(ngrp + 1, 0);
      for (int g = 0; g < ngrp; g++) {
        wstart[g] = (int)wown.size();
        for (int iv : by[g]) {
          int src = 0;
          if (per) { const int sh = iv / per, j = iv % per; src = (d.ksplit + (j >> 3)) * d.pitch0 + sh * 8 + (j & 7); }
          wown.push_back(src | (colpos[iv] << 16));
          wiv.push_back(iv);
        }
      }
      wstart[ngrp] = (int)wown.size();
      if ((rc = upload(p, wown, &d.wown)) || (rc = upload(p, wiv, &d.wown_iv)) || (rc = upload(p, wstart, &d.wown_start))) { qb200_plan_destroy(p); return rc; }
      p->smem_plane += (size_t)d.ncolpos_c * 16;
    }
    if ((rc = plane_opt_in(p))) { qb200_plan_destroy(p); return rc; }
    // tensor-memory kernel of the compiled shape: positions in a buffer that holds the kept rows only
    p->plane_t = false;
    if (plane_t_wanted(p) && d.nvec <= 65535) {
      int xs, xk;
      plane_t_xrange(&xs, &xk);
      if ((rc = build_kept_row_tables(plane_t_pitch(), xs, xk)) || (rc = plane_t_setup(p))) { qb200_plan_destroy(p); return rc; }
    }
  }
  // planes whose kept rows fit shared memory once (si54p 126 x 126): the whole xy stage in one kernel (k_plane_f)
  p->plane_f = false;
  if (!p->fused) {
    int hmax = 0;
    for (int r = 0; r < nrods; r++) hmax = std::max(hmax, std::abs(rod_h[r]));
    if (plane_f_wanted(p, hmax)) {
      int xs, xk;
      plane_f_xrange(&xs, &xk);
      if ((rc = build_kept_row_tables(plane_f_pitch(), xs, xk)) || (rc = plane_f_setup(p))) { qb200_plan_destroy(p); return rc; }
    }
  }
  // fused path: batches of up to ~1100 MgO216 states -- whole blocks go through in one batch, so the persistent CTAs of
  // the plane and z-column kernels pay their wave quantisation and prologue once (r1l sweep: 256 MiB 27.66 ms/step, 1 GiB 26.68,
  // 4 GiB 26.33); the scratch is allocated for the units actually batched, never for the whole budget
  p->ws_bytes = p->fused ? (4ll << 30) : (8ll << 30);
  if (const char* e = getenv("QB200_WORKSPACE_BYTES")) p->ws_bytes = atoll(e);
  configure_batch(p);
  *out = p;
  return QB200_OK;
}

extern "C" int qb200_plan_destroy(qb200_plan* p)
{
  if (!p) return QB200_OK;
  cudaSetDevice(p->device);
  for (void* q : p->owned) cudaFree(q);
  for (cudaEvent_t e : p->evs) cudaEventDestroy(e);
  if (p->s_in) cudaStreamDestroy(p->s_in);
  if (p->s_out) cudaStreamDestroy(p->s_out);
  for (double* q : { p->zt, p->w, p->rho_part, p->fac_dev, p->st_c, p->st_cp, p->st_v, p->st_f, p->st_kpg2, p->ex_a, p->ex_b, p->ex_c2, p->vh })
    if (q) cudaFree(q);
  delete p;
  return QB200_OK;
}

extern "C" int qb200_plan_set_stream(qb200_plan* p, void* s)
{
  if (!p) return QB200_EINVAL;
  p->stream = (cudaStream_t)s;
  return QB200_OK;
}

extern "C" int qb200_plan_set_coefficient_tag(qb200_plan* p, long long tag)
{
  if (!p) return QB200_EINVAL;
  p->next_tag = tag;
  return QB200_OK;
}

extern "C" int qb200_plan_set_workspace(qb200_plan* p, long long bytes)
{
  if (!p || bytes < 0) return QB200_EINVAL;
  p->ws_bytes = bytes;
  return configure_batch(p);
}

extern "C" long long qb200_plan_query(const qb200_plan* p, int what)
{
  if (!p) return -1;
  switch (what) {
    case 0: return p->d.np0;
    case 1: return p->d.np1;
    case 2: return p->d.np2;
    case 3: return p->d.nvec;
    case 4: return p->d.ntrans0;
    case 5: return p->d.ngw;
    case 6: return p->d.is_real;
    case 7: return p->fused ? 1 : 0;
    case 8: return p->batch;
    case 9: return p->launches;
    case 10: return p->static_shape;
    case 11: return p->z2 ? 1 : 0;
    case 14: return p->split2 ? 1 : 0;
    case 15: return p->split_static;
    case 16: return p->z_static;
    case 17: return p->plane_t ? 1 : 0;
    case 18: return p->zcol_t ? 1 : 0;
    case 19: return p->ycols_t;
    case 20: return p->plane_f ? 1 : 0;
    case 12: return p->d.zb_cb;
    case 13: return p->d.zf_cb;
    default: return -1;
  }
}

// ------------------------------------------------------------------------------------------------ launch helpers
#define QB_LAUNCH_CHECK(p) do { (p)->launches++; cudaError_t e__ = cudaGetLastError(); \
    if (e__ != cudaSuccess) return qb200::cuda_fail(e__, "kernel launch", __FILE__, __LINE__); } while (0)

static int nzblocks(const qb200_plan* p) { return (p->d.nrods + p->d.rb - 1) / p->d.rb; }

// persistent grid of the v2 z kernels: (rod blocks, G) filling the resident-CTA slots
static dim3 z2_grid(const qb200_plan* p, int rb, int slots, int nunits)
{
  const int nzb = (p->d.nrods + rb - 1) / rb;
  return dim3(nzb, std::max(1, std::min(nunits, slots / nzb)));
}

static int launch_zbwd(qb200_plan* p, int mode, const double* c, size_t ldc, int nunits)
{
  dim3 g(nzblocks(p), nunits);
  prof_begin(0, p->stream);
  if (p->zcol_t && mode == MODE_SINGLE) {
    const int rc = launch_zbwd_t(p, c, ldc, nunits);
    prof_end(p->stream);
    if (rc) return rc;
    p->launches++;
    return QB200_OK;
  }
  if (p->z2) {
    const dim3 g2 = z2_grid(p, p->d.zb_rb, p->zslots_b[mode], nunits);
    const int zt = p->z_threads;
    if (mode == MODE_PAIR) k_zcol_bwd2<MODE_PAIR, DynZ><<<g2, zt, p->smem_zb[1], p->stream>>>(p->d, (const cplx*)c, ldc, (cplx*)p->zt, nunits);
    else if (p->z_static == 1 && zt == 128) k_zcol_bwd2<MODE_SINGLE, ZbMgO216n><<<g2, zt, p->smem_zb[0], p->stream>>>(p->d, (const cplx*)c, ldc, (cplx*)p->zt, nunits);
    else if (p->z_static == 1) k_zcol_bwd2<MODE_SINGLE, ZbMgO216><<<g2, zt, p->smem_zb[0], p->stream>>>(p->d, (const cplx*)c, ldc, (cplx*)p->zt, nunits);
    else k_zcol_bwd2<MODE_SINGLE, DynZ><<<g2, zt, p->smem_zb[0], p->stream>>>(p->d, (const cplx*)c, ldc, (cplx*)p->zt, nunits);
  }
  else if (mode == MODE_PAIR) k_zcol_bwd<MODE_PAIR><<<g, 256, p->smem_z, p->stream>>>(p->d, (const cplx*)c, ldc, (cplx*)p->zt);
  else k_zcol_bwd<MODE_SINGLE><<<g, 256, p->smem_z, p->stream>>>(p->d, (const cplx*)c, ldc, (cplx*)p->zt);
  prof_end(p->stream);
  QB_LAUNCH_CHECK(p);
  return QB200_OK;
}

static int launch_zfwd(qb200_plan* p, int mode, double* out, size_t ldc, int nunits, int accumulate, const double* kpg2,
                       const double* cin)
{
  dim3 g(nzblocks(p), nunits);
  const double scale = 1.0 / ((double)p->d.np0 * p->d.np1 * p->d.np2);
  prof_begin(2, p->stream);
  if (p->zcol_t && mode == MODE_SINGLE) {
    const int rc = launch_zfwd_t(p, out, ldc, nunits, accumulate, kpg2, cin, scale);
    prof_end(p->stream);
    if (rc) return rc;
    p->launches++;
    return QB200_OK;
  }
  if (p->z2) {
    const dim3 g2 = z2_grid(p, p->d.zf_rb, p->zslots_f[mode], nunits);
    const int zt = p->z_threads;
    if (mode == MODE_PAIR)
      k_zcol_fwd2<MODE_PAIR, DynZ><<<g2, zt, p->smem_zf[1], p->stream>>>(p->d, (const cplx*)p->zt, (cplx*)out, ldc, accumulate, kpg2, (const cplx*)cin, scale, nunits);
    else if (p->z_static == 1 && zt == 128)
      k_zcol_fwd2<MODE_SINGLE, ZfMgO216n><<<g2, zt, p->smem_zf[0], p->stream>>>(p->d, (const cplx*)p->zt, (cplx*)out, ldc, accumulate, kpg2, (const cplx*)cin, scale, nunits);
    else if (p->z_static == 1)
      k_zcol_fwd2<MODE_SINGLE, ZfMgO216><<<g2, zt, p->smem_zf[0], p->stream>>>(p->d, (const cplx*)p->zt, (cplx*)out, ldc, accumulate, kpg2, (const cplx*)cin, scale, nunits);
    else
      k_zcol_fwd2<MODE_SINGLE, DynZ><<<g2, zt, p->smem_zf[0], p->stream>>>(p->d, (const cplx*)p->zt, (cplx*)out, ldc, accumulate, kpg2, (const cplx*)cin, scale, nunits);
  }
  else if (mode == MODE_PAIR)
    k_zcol_fwd<MODE_PAIR><<<g, 256, p->smem_z, p->stream>>>(p->d, (const cplx*)p->zt, (cplx*)out, ldc, accumulate, kpg2, (const cplx*)cin, scale);
  else
    k_zcol_fwd<MODE_SINGLE><<<g, 256, p->smem_z, p->stream>>>(p->d, (const cplx*)p->zt, (cplx*)out, ldc, accumulate, kpg2, (const cplx*)cin, scale);
  prof_end(p->stream);
  QB_LAUNCH_CHECK(p);
  return QB200_OK;
}

// the three kernels of the split xy stage for one engine / compiled shape
template <int OP, class SH>
static int launch_split(qb200_plan* p, dim3 gr, dim3 gy, const double* v, double* f, const double* fac, int nunits, int zero_imag)
{
  const DevPlan& d = p->d;
  if (OP != OP_FWD) {
    k_xrows2<+1, SH><<<gr, 256, p->smem_xr[0], p->stream>>>(d, (cplx*)p->zt, (cplx*)p->w, p->xr_rowb, p->xr_smax, nunits);
    QB_LAUNCH_CHECK(p);
  }
  if (p->ycols_t && (OP == OP_HPSI || OP == OP_DENSITY)) {
    const int rc = launch_ycols_t(p, OP, v, fac, nunits, zero_imag);
    if (rc) return rc;
    p->launches++;
  } else {
    k_ycols2<OP, SH><<<gy, 256, p->smem_yc[OP], p->stream>>>(d, (cplx*)p->w, v, (cplx*)f, p->rho_part, fac, nunits, zero_imag);
    QB_LAUNCH_CHECK(p);
  }
  if (OP == OP_HPSI || OP == OP_FWD) {
    k_xrows2<-1, SH><<<gr, 256, p->smem_xr[1], p->stream>>>(d, (cplx*)p->zt, (cplx*)p->w, p->xr_rowb, p->xr_smax, nunits);
    QB_LAUNCH_CHECK(p);
  }
  return QB200_OK;
}

// the xy stage for `nunits` units whose column data sits in p->zt
template <int OP>
static int launch_xy(qb200_plan* p, int nunits, const double* v, double* f, const double* fac, int ngroups, int zero_imag)
{
  const DevPlan& d = p->d;
  const int upg = (OP == OP_DENSITY) ? (nunits + ngroups - 1) / ngroups : 1;
  const int ng = (OP == OP_DENSITY) ? ngroups : nunits;
  if (p->fused) {
    dim3 g(d.np2, std::max(1, std::min(ngroups, nunits)));
    if (OP == OP_DENSITY) prof_begin(8, p->stream);      // (category 8: the density launches of category 1)
    prof_begin(1, p->stream);
    const int rc = launch_plane(p, OP, g, v, f, fac, nunits, zero_imag);
    prof_end(p->stream);
    if (rc) return rc;
    p->launches++;
    return QB200_OK;
  }
  if (p->plane_f && !zero_imag && (OP == OP_HPSI || OP == OP_DENSITY)) {
    dim3 g(d.np2, std::max(1, std::min(ngroups, nunits)));
    if (OP == OP_DENSITY) prof_begin(8, p->stream);
    prof_begin(1, p->stream);
    const int rc = launch_plane_f(p, OP, g, v, fac, nunits);
    prof_end(p->stream);
    if (rc) return rc;
    p->launches++;
    return QB200_OK;
  }
  if (p->split2) {
    // persistent CTAs per (block, plane); G spreads the units of the batch only when the planes alone cannot fill the SMs
    const int nrb = (d.nkeep + p->xr_rowb - 1) / p->xr_rowb, nxb = (d.np0 + d.xb - 1) / d.xb;
    auto spread = [&](long ctas) { return (int)std::max(1l, std::min<long>(nunits, (4l * p->nsm + ctas - 1) / ctas)); };
    const dim3 gr(nrb, d.np2, spread((long)nrb * d.np2));
    const dim3 gy(nxb, d.np2, OP == OP_DENSITY ? std::max(1, std::min(ngroups, nunits)) : spread((long)nxb * d.np2));
    prof_begin(1, p->stream);
    int rc;
    switch (p->split_static) {
      case 1: rc = launch_split<OP, ShapeAu992>(p, gr, gy, v, f, fac, nunits, zero_imag); break;
      case 2: rc = launch_split<OP, ShapeSi54p>(p, gr, gy, v, f, fac, nunits, zero_imag); break;
      default: rc = launch_split<OP, DynSplit>(p, gr, gy, v, f, fac, nunits, zero_imag); break;
    }
    prof_end(p->stream);
    return rc;
  }
  const int rowb = (int)((p->smem_rows / 16 - d.np0) / d.pitch0);
  dim3 gr((d.nkeep + rowb - 1) / rowb, d.np2, nunits);
  dim3 gy((d.np0 + d.xb - 1) / d.xb, d.np2, ng);
  prof_begin(1, p->stream);
  if (OP != OP_FWD) {
    k_xrows<+1><<<gr, 256, p->smem_rows, p->stream>>>(d, (cplx*)p->zt, (cplx*)p->w, rowb);
    QB_LAUNCH_CHECK(p);
  }
  k_ycols<OP><<<gy, 256, p->smem_ycol, p->stream>>>(d, (cplx*)p->w, v, (cplx*)f, p->rho_part, fac, nunits, upg, zero_imag);
  QB_LAUNCH_CHECK(p);
  if (OP == OP_HPSI || OP == OP_FWD) {
    k_xrows<-1><<<gr, 256, p->smem_rows, p->stream>>>(d, (cplx*)p->zt, (cplx*)p->w, rowb);
    QB_LAUNCH_CHECK(p);
  }
  prof_end(p->stream);
  return QB200_OK;
}

// host<->device staging
struct Staged {
  const double* dev; bool staged;
};
static int stage_in(qb200_plan* p, const double* ptr, size_t elems, double** buf, size_t* cap, const double** dev)
{
  if (!ptr) { *dev = nullptr; return QB200_OK; }
  if (is_device_ptr(ptr)) { *dev = ptr; return QB200_OK; }
  if (buf == &p->st_c) p->res_ptr = nullptr;            // st_c no longer holds a tagged coefficient block
  int rc = ensure(buf, cap, elems);
  if (rc) return rc;
  QB_CUDA(cudaMemcpyAsync(*buf, ptr, elems * sizeof(double), cudaMemcpyHostToDevice, p->stream));
  *dev = *buf;
  return QB200_OK;
}

// ------------------------------------------------------------------------------------------------ single transforms
static int fft_backward_impl(qb200_plan* p, const double* c1, const double* c2, double* f)
{
  if (!p || !c1 || !f) { set_error("qb200_fft_backward: bad argument"); return QB200_EINVAL; }
  if (c2 && !p->d.is_real) { set_error("qb200_fft_backward_pair: basis is not real (FourierTransform.cc:1688 asserts)"); return QB200_EINVAL; }
  QB_CUDA(cudaSetDevice(p->device));
  const DevPlan& d = p->d;
  const size_t N = (size_t)d.np0 * d.np1 * d.np2;
  int rc = ensure_work(p, 1);
  if (rc) return rc;
  // coefficients: gather c1 (and c2) into one ldc = ngw block of 1 or 2 columns on the device
  p->res_ptr = nullptr;
  rc = ensure(&p->st_c, &p->st_c_cap, 4 * (size_t)d.ngw);
  if (rc) return rc;
  const cudaMemcpyKind any = cudaMemcpyDefault;
  QB_CUDA(cudaMemcpyAsync(p->st_c, c1, 16 * (size_t)d.ngw, any, p->stream));
  if (c2) QB_CUDA(cudaMemcpyAsync(p->st_c + 2 * (size_t)d.ngw, c2, 16 * (size_t)d.ngw, any, p->stream));
  double* fdev = f;
  const bool fhost = !is_device_ptr(f);
  if (fhost) { rc = ensure(&p->st_f, &p->st_f_cap, 2 * N); if (rc) return rc; fdev = p->st_f; }
  if ((rc = launch_zbwd(p, c2 ? MODE_PAIR : MODE_SINGLE, p->st_c, d.ngw, 1))) return rc;
  if ((rc = launch_xy<OP_BWD>(p, 1, nullptr, fdev, nullptr, 1, 0))) return rc;
  if (fhost) QB_CUDA(cudaMemcpyAsync(f, fdev, 16 * N, cudaMemcpyDeviceToHost, p->stream));
  QB_CUDA(cudaStreamSynchronize(p->stream));
  return QB200_OK;
}

static int fft_forward_impl(qb200_plan* p, double* f, double* c1, double* c2)
{
  if (!p || !c1 || !f) { set_error("qb200_fft_forward: bad argument"); return QB200_EINVAL; }
  if (c2 && !p->d.is_real) { set_error("qb200_fft_forward_pair: basis is not real (FourierTransform.cc:1727 asserts)"); return QB200_EINVAL; }
  QB_CUDA(cudaSetDevice(p->device));
  const DevPlan& d = p->d;
  const size_t N = (size_t)d.np0 * d.np1 * d.np2;
  int rc = ensure_work(p, 1);
  if (rc) return rc;
  const double* fdev;
  if ((rc = stage_in(p, f, 2 * N, &p->st_f, &p->st_f_cap, &fdev))) return rc;
  p->res_ptr = nullptr;
  rc = ensure(&p->st_c, &p->st_c_cap, 4 * (size_t)d.ngw);
  if (rc) return rc;
  if ((rc = launch_xy<OP_FWD>(p, 1, nullptr, (double*)fdev, nullptr, 1, 0))) return rc;
  if ((rc = launch_zfwd(p, c2 ? MODE_PAIR : MODE_SINGLE, p->st_c, d.ngw, 1, 0, nullptr, nullptr))) return rc;
  QB_CUDA(cudaMemcpyAsync(c1, p->st_c, 16 * (size_t)d.ngw, cudaMemcpyDefault, p->stream));
  if (c2) QB_CUDA(cudaMemcpyAsync(c2, p->st_c + 2 * (size_t)d.ngw, 16 * (size_t)d.ngw, cudaMemcpyDefault, p->stream));
  QB_CUDA(cudaStreamSynchronize(p->stream));
  return QB200_OK;
}

extern "C" int qb200_fft_backward(qb200_plan* p, const double* c, double* f) { return fft_backward_impl(p, c, nullptr, f); }
extern "C" int qb200_fft_backward_pair(qb200_plan* p, const double* c1, const double* c2, double* f)
{
  if (!c2) { set_error("qb200_fft_backward_pair: bad argument"); return QB200_EINVAL; }
  return fft_backward_impl(p, c1, c2, f);
}
extern "C" int qb200_fft_forward(qb200_plan* p, double* f, double* c) { return fft_forward_impl(p, f, c, nullptr); }
extern "C" int qb200_fft_forward_pair(qb200_plan* p, double* f, double* c1, double* c2)
{
  if (!c2) { set_error("qb200_fft_forward_pair: bad argument"); return QB200_EINVAL; }
  return fft_forward_impl(p, f, c1, c2);
}

// ------------------------------------------------------------------------------------------------ rs_mul_add
// device pointers only (also used by hpsi.cu)
int qb200_rs_mul_add_dev(qb200_plan* p, int ldc, int nst, const double* c, const double* v, const double* kpg2, double* cp)
{
  const DevPlan& d = p->d;
  int rc;
  auto run = [&](int first_state, int nunits, int mode, int zero_imag) -> int {
    const int spu = mode == MODE_PAIR ? 2 : 1;
    for (int b0 = 0; b0 < nunits;) {
      int nb, G;
      plan_split(p, nunits - b0, 64, &nb, &G);
      int r = ensure_work(p, nb);
      if (r) return r;
      const size_t off = 2 * (size_t)(first_state + b0 * spu) * ldc;
      if ((r = launch_zbwd(p, mode, c + off, ldc, nb))) return r;
      if ((r = launch_xy<OP_HPSI>(p, nb, v, nullptr, nullptr, G, zero_imag))) return r;
      if ((r = launch_zfwd(p, mode, cp + off, ldc, nb, 1, kpg2, kpg2 ? c + off : nullptr))) return r;
      b0 += nb;
    }
    return QB200_OK;
  };
  if (d.is_real) {                                   // SlaterDet.cc:987-1023: local pairs, then the odd tail with Im := 0
    const int npair = nst / 2;
    if (npair && (rc = run(0, npair, MODE_PAIR, 0))) return rc;
    if (nst % 2 && (rc = run(nst - 1, 1, MODE_SINGLE, 1))) return rc;
  } else {
    if ((rc = run(0, nst, MODE_SINGLE, 0))) return rc;  // SlaterDet.cc:1027-1037
  }
  return QB200_OK;
}

extern "C" int qb200_rs_mul_add(qb200_plan* p, int ldc, int nst, const double* c, const double* v, const double* kpg2,
                                double* cp)
{
  if (!p || !c || !v || !cp || nst < 0 || ldc < p->d.ngw) { set_error("qb200_rs_mul_add: bad argument"); return QB200_EINVAL; }
  if (nst == 0) return QB200_OK;
  QB_CUDA(cudaSetDevice(p->device));
  const DevPlan& d = p->d;
  const size_t N = (size_t)d.np0 * d.np1 * d.np2, blk = 2 * (size_t)ldc * nst;
  const double *cd, *vd, *kd, *cpd_c;
  int rc;
  if ((rc = stage_in(p, c, blk, &p->st_c, &p->st_c_cap, &cd))) return rc;
  if ((rc = stage_in(p, v, N, &p->st_v, &p->st_v_cap, &vd))) return rc;
  if ((rc = stage_in(p, kpg2, d.ngw, &p->st_kpg2, &p->st_kpg2_cap, &kd))) return rc;
  if ((rc = stage_in(p, cp, blk, &p->st_cp, &p->st_cp_cap, &cpd_c))) return rc;
  double* cpd = const_cast<double*>(cpd_c);
  if ((rc = qb200_rs_mul_add_dev(p, ldc, nst, cd, vd, kd, cpd))) return rc;
  if (cpd != cp) {
    QB_CUDA(cudaMemcpyAsync(cp, cpd, blk * sizeof(double), cudaMemcpyDeviceToHost, p->stream));
    QB_CUDA(cudaStreamSynchronize(p->stream));
  }
  return QB200_OK;
}

// ------------------------------------------------------------------------------------------------ compute_density
extern "C" int qb200_compute_density(qb200_plan* p, int ldc, int nst, const double* c, const double* fac, double* rho)
{
  if (!p || !c || !fac || !rho || nst < 0 || ldc < p->d.ngw) { set_error("qb200_compute_density: bad argument"); return QB200_EINVAL; }
  if (nst == 0) return QB200_OK;
  QB_CUDA(cudaSetDevice(p->device));
  const DevPlan& d = p->d;
  const size_t N = (size_t)d.np0 * d.np1 * d.np2, blk = 2 * (size_t)ldc * nst;
  const double *cd = c, *rd_c;
  int rc;
  // a host coefficient block is uploaded batch by batch on the copy stream while earlier batches are transformed
  // (or not at all when the caller tagged it as unchanged since the upload that left it in st_c)
  const bool chost = !is_device_ptr(c);
  const bool upload_c = chost && !plan_resident(p, c, ldc, nst);
  if (chost) {
    if ((rc = ensure(&p->st_c, &p->st_c_cap, blk))) return rc;
    cd = p->st_c;
    if (upload_c) { p->res_ptr = nullptr; if ((rc = plan_copy_streams(p))) return rc; }
  }
  // a host rho (the caller's running sum: rho += ...) is uploaded on the copy stream under the transforms -- only the final
  // reduction reads it -- after the coefficient blocks, which are needed first
  const bool rhost = !is_device_ptr(rho);
  cudaEvent_t ev_rho = nullptr;
  if (rhost) {
    if ((rc = ensure(&p->st_v, &p->st_v_cap, N)) || (rc = plan_copy_streams(p))) return rc;
    rd_c = p->st_v;
  } else rd_c = rho;
  double* rd = const_cast<double*>(rd_c);
  if ((rc = ensure(&p->fac_dev, &p->fac_cap, nst))) return rc;
  // Real bases at Gamma: two states per transform, psi_1 + i psi_2 (the packing of FourierTransform.cc:555-581 that
  // SlaterDet.cc:858-885 applies to the density: rho += fac1 Re^2 + fac2 Im^2), the odd last state alone (:886-903).
  // fac_dev then holds [first weights | second weights | the odd state's weight].  QB200_DENSITY_PAIRS=0: one state per transform.
  const bool pairs_on = !(getenv("QB200_DENSITY_PAIRS") && atoi(getenv("QB200_DENSITY_PAIRS")) == 0);
  int npair = (d.is_real && pairs_on) ? nst / 2 : 0;
  if (npair) {
    // The packing needs what the reference guarantees for real bases: Im c_n(G=0) = 0 (SlaterDet.cc:2776-2779).  A state whose
    // G = 0 coefficient carries an imaginary part would leak it into its partner's density at first order, while the
    // one-state transform the reference runs (:906-924) sees it at second order only -- such a block takes the reference's branch.
    bool heads_real = true;
    if (chost) {
      for (int n = 0; n < nst && heads_real; n++) {
        const double* col = c + 2 * (size_t)n * ldc;
        double ref = 0.0;
        for (int i = 0; i < std::min(64, d.ngw); i++) ref = std::max(ref, std::max(std::fabs(col[2 * i]), std::fabs(col[2 * i + 1])));
        if (std::fabs(col[1]) > 1e-13 * ref) heads_real = false;
      }
    } else {
      int* flag = reinterpret_cast<int*>(p->fac_dev);          // (fac_dev is filled after this check)
      int h = 0;
      QB_CUDA(cudaMemsetAsync(flag, 0, sizeof(int), p->stream));
      k_gamma_heads<<<(nst + 127) / 128, 128, 0, p->stream>>>((const cplx*)c, (size_t)ldc, d.ngw, nst, flag);
      QB_LAUNCH_CHECK(p);
      QB_CUDA(cudaMemcpyAsync(&h, flag, sizeof(int), cudaMemcpyDeviceToHost, p->stream));
      QB_CUDA(cudaStreamSynchronize(p->stream));
      heads_real = h == 0;
    }
    if (!heads_real) npair = 0;
  }
  if (npair && !is_device_ptr(fac)) {
    p->fac_host.resize(nst);
    for (int u = 0; u < npair; u++) { p->fac_host[u] = fac[2 * u]; p->fac_host[npair + u] = fac[2 * u + 1]; }
    if (nst % 2) p->fac_host[2 * npair] = fac[nst - 1];
    QB_CUDA(cudaMemcpyAsync(p->fac_dev, p->fac_host.data(), nst * sizeof(double), cudaMemcpyHostToDevice, p->stream));
  } else if (npair) {
    QB_CUDA(cudaMemcpy2DAsync(p->fac_dev, sizeof(double), fac, 2 * sizeof(double), sizeof(double), npair, cudaMemcpyDeviceToDevice, p->stream));
    QB_CUDA(cudaMemcpy2DAsync(p->fac_dev + npair, sizeof(double), fac + 1, 2 * sizeof(double), sizeof(double), npair, cudaMemcpyDeviceToDevice, p->stream));
    if (nst % 2) QB_CUDA(cudaMemcpyAsync(p->fac_dev + 2 * npair, fac + nst - 1, sizeof(double), cudaMemcpyDeviceToDevice, p->stream));
  } else {
    QB_CUDA(cudaMemcpyAsync(p->fac_dev, fac, nst * sizeof(double), cudaMemcpyDefault, p->stream));
  }
  // batches of units (a unit = one state or one pair); groups: exclusive owners of a partial density each; enough CTAs to
  // fill the machine
  struct Batch { int state0, nunits, spu, facoff, G; };
  std::vector<Batch> batches;
  const int maxG = 16;
  int ngroups = 1;
  auto add_batches = [&](int state0, int nunits, int spu, int facoff) {
    for (int b0 = 0; b0 < nunits;) {
      int nb, G;
      plan_split(p, nunits - b0, maxG, &nb, &G);
      ngroups = std::max(ngroups, G);
      batches.push_back({state0 + b0 * spu, nb, spu, facoff + b0, G});
      b0 += nb;
    }
  };
  if (npair) {
    add_batches(0, npair, 2, 0);
    if (nst % 2) add_batches(nst - 1, 1, 1, 2 * npair);
  } else {
    add_batches(0, nst, 1, 0);
  }
  if (upload_c) {
    cudaEvent_t ev;
    if ((rc = plan_event(p, 0, &ev))) return rc;
    QB_CUDA(cudaEventRecord(ev, p->stream));              // st_c may still be read by earlier work on the plan's stream
    QB_CUDA(cudaStreamWaitEvent(p->s_in, ev, 0));
    for (size_t ib = 0; ib < batches.size(); ib++) {
      const Batch& b = batches[ib];
      const size_t off = 2 * (size_t)b.state0 * ldc, cnt = 2 * (size_t)b.nunits * b.spu * ldc;
      QB_CUDA(cudaMemcpyAsync(p->st_c + off, c + off, cnt * sizeof(double), cudaMemcpyHostToDevice, p->s_in));
      if ((rc = plan_event(p, 1 + (int)ib, &ev))) return rc;
      QB_CUDA(cudaEventRecord(ev, p->s_in));
    }
  }
  if (rhost) {
    cudaEvent_t e0;
    if ((rc = plan_event(p, 102, &e0)) || (rc = plan_event(p, 103, &ev_rho))) return rc;
    QB_CUDA(cudaEventRecord(e0, p->stream));              // st_v may still be read by earlier work on the plan's stream
    QB_CUDA(cudaStreamWaitEvent(p->s_in, e0, 0));
    QB_CUDA(cudaMemcpyAsync(p->st_v, rho, N * sizeof(double), cudaMemcpyHostToDevice, p->s_in));
    QB_CUDA(cudaEventRecord(ev_rho, p->s_in));
  }
  if ((rc = ensure(&p->rho_part, &p->rho_part_elems, (size_t)ngroups * N))) return rc;
  QB_CUDA(cudaMemsetAsync(p->rho_part, 0, (size_t)ngroups * N * sizeof(double), p->stream));
  for (size_t ib = 0; ib < batches.size(); ib++) {
    const Batch& b = batches[ib];
    if ((rc = ensure_work(p, b.nunits))) return rc;
    if (upload_c) QB_CUDA(cudaStreamWaitEvent(p->stream, p->evs[1 + ib], 0));
    if ((rc = launch_zbwd(p, b.spu == 2 ? MODE_PAIR : MODE_SINGLE, cd + 2 * (size_t)b.state0 * ldc, ldc, b.nunits))) return rc;
    p->d.fac2off = b.spu == 2 ? npair : 0;               // (the plan descriptor travels by value with each launch)
    rc = launch_xy<OP_DENSITY>(p, b.nunits, nullptr, nullptr, p->fac_dev + b.facoff, (p->fused || p->plane_f) ? b.G : 1, 0);
    p->d.fac2off = 0;
    if (rc) return rc;
  }
  if (ev_rho) QB_CUDA(cudaStreamWaitEvent(p->stream, ev_rho, 0));
  prof_begin(6, p->stream);
  k_rho_reduce<<<std::min<size_t>((N + 255) / 256, 148 * 8), 256, 0, p->stream>>>(rd, p->rho_part, N, ngroups);
  prof_end(p->stream);
  QB_LAUNCH_CHECK(p);
  if (rd != rho) QB_CUDA(cudaMemcpyAsync(rho, rd, N * sizeof(double), cudaMemcpyDeviceToHost, p->stream));
  if (rd != rho || chost) QB_CUDA(cudaStreamSynchronize(p->stream));
  if (upload_c) plan_mark_resident(p, c, ldc, nst);
  return QB200_OK;
}

// ------------------------------------------------------------------------------------------------ update_density tail
// ChargeDensity::update_density after the row sum (ChargeDensity.cc:516-551): nelectrons = sum(rho)*omega/N,
// rhotmp = omega*rho, rhog = vft->forward(rhotmp).  `pv` is the plan of the DENSITY basis (vbasis_, ChargeDensity.cc:77-81).
extern "C" int qb200_density_finish(qb200_plan* pv, const double* rho, double omega, double* rhog, double* nelectrons)
{
  if (!pv || !rho || !rhog || !(omega > 0.0)) { set_error("qb200_density_finish: bad argument"); return QB200_EINVAL; }
  QB_CUDA(cudaSetDevice(pv->device));
  const DevPlan& d = pv->d;
  const size_t N = (size_t)d.np0 * d.np1 * d.np2;
  const double* rd;
  int rc;
  if ((rc = stage_in(pv, rho, N, &pv->st_v, &pv->st_v_cap, &rd))) return rc;
  const int nblk = 148 * 4;
  if ((rc = ensure(&pv->st_f, &pv->st_f_cap, 2 * N + nblk + 2))) return rc;
  double* f = pv->st_f;
  double* bs = f + 2 * N;
  k_rho_expand<<<nblk, 256, 0, pv->stream>>>(rd, N, omega, (cplx*)f, bs);
  QB_LAUNCH_CHECK(pv);
  k_sum_fixed<<<1, 32, 0, pv->stream>>>(bs, nblk, omega / (double)N, bs + nblk);
  QB_LAUNCH_CHECK(pv);
  double nel = 0.0;
  QB_CUDA(cudaMemcpyAsync(&nel, bs + nblk, sizeof(double), cudaMemcpyDeviceToHost, pv->stream));
  if ((rc = fft_forward_impl(pv, f, rhog, nullptr))) return rc;          // synchronises the stream
  if (nelectrons) *nelectrons = nel;
  return QB200_OK;
}

#include "vhxc.cuh"
