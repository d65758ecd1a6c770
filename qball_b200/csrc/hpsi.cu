// qball_b200/csrc/hpsi.cu -- the H psi column block of EnergyFunctional::energy(compute_hpsi = true)
// (/root/reference/src/qball/EnergyFunctional.cc:1142-1153 clear, :1500 nonlocal, :1675-1690 kinetic, :1695 local),
// in the reference's order, on one stream, with host blocks staged in and out when host pointers are given.
#include "qb200_internal.h"
#include <algorithm>
#include <cstdlib>

int qb200_rs_mul_add_dev(qb200_plan* p, int ldc, int nst, const double* c, const double* v, const double* kpg2, double* cp);
int qb200_nl_energy_dev(qb200_nl* nl, int ldc, int nst, const double* c, const double* occ_host, int compute_hpsi, double* cp, int cont);
int qb200_nl_chunks(const qb200_nl* nl, int nst);
int qb200_nl_projectors(const qb200_nl* nl);
double* qb200_nl_enl_dev(qb200_nl* nl);
cudaStream_t qb200_nl_swap_stream(qb200_nl* nl, cudaStream_t s);

using namespace qb200;

static int ensure_buf(double** buf, size_t* cap, size_t elems)
{
  if (*cap >= elems && *buf) return QB200_OK;
  if (*buf) { cudaFree(*buf); *buf = nullptr; *cap = 0; }
  QB_CUDA(cudaMalloc((void**)buf, std::max<size_t>(elems, 1) * sizeof(double)));
  *cap = elems;
  return QB200_OK;
}

// states per pipeline slice of the host-pointer path: whole GEMM tiles (64 states), an even count (real bases pair
// local states (n, n+1), SlaterDet.cc:987), at most 16 slices -- small slices keep the exposed first upload and last
// download short; one slice when the projector sweep is chunked (each slice would regenerate every anl chunk)
static int hpsi_block_states(int nst, bool single)
{
  int unit = 64;
  if (const char* e = getenv("QB200_HOST_SLICE")) { const int v = atoi(e); if (v >= 2 && v % 2 == 0) unit = v; }
  if (single || nst <= unit) return nst;
  const int nblk = std::min(16, (nst + unit - 1) / unit);
  const int per = (nst + nblk - 1) / nblk;
  return (per + unit - 1) / unit * unit;
}

extern "C" int qb200_hpsi(qb200_plan* p, qb200_nl* nl, int ldc, int nst, const double* c, const double* occ, const double* v,
                          const double* kpg2, double* hpsi, double* enl)
{
  if (!p || !c || !v || !hpsi || nst < 0 || ldc < p->d.ngw || (nl && !occ)) { set_error("qb200_hpsi: bad argument"); return QB200_EINVAL; }
  if (enl) *enl = 0.0;
  if (nst == 0) return QB200_OK;
  QB_CUDA(cudaSetDevice(p->device));
  const DevPlan& d = p->d;
  const size_t N = (size_t)d.np0 * d.np1 * d.np2, blk = 2 * (size_t)ldc * nst;
  int rc;
  const double *cd = c, *vd = v, *kd = kpg2;
  double* od = hpsi;
  const bool chost = !is_device_ptr(c), ohost = !is_device_ptr(hpsi);
  const bool upload_c = chost && !plan_resident(p, c, ldc, nst);
  if (chost) {
    if ((rc = ensure_buf(&p->st_c, &p->st_c_cap, blk))) return rc;
    cd = p->st_c;
    if (upload_c) p->res_ptr = nullptr;
  }
  cudaEvent_t ev_v = nullptr;
  if (!is_device_ptr(v)) {
    // v(r) from the host rides on the copy stream under the projector contraction (the first term of H psi does not read it);
    // the local term of the first block waits for it
    if ((rc = ensure_buf(&p->st_v, &p->st_v_cap, N)) || (rc = plan_copy_streams(p))) return rc;
    cudaEvent_t e0;
    if ((rc = plan_event(p, 100, &e0)) || (rc = plan_event(p, 101, &ev_v))) return rc;
    QB_CUDA(cudaEventRecord(e0, p->stream));              // st_v may still be read by earlier work on the plan's stream
    QB_CUDA(cudaStreamWaitEvent(p->s_in, e0, 0));
    QB_CUDA(cudaMemcpyAsync(p->st_v, v, N * sizeof(double), cudaMemcpyHostToDevice, p->s_in));
    QB_CUDA(cudaEventRecord(ev_v, p->s_in));
    vd = p->st_v;
  }
  if (kpg2 && !is_device_ptr(kpg2)) {
    if ((rc = ensure_buf(&p->st_kpg2, &p->st_kpg2_cap, d.ngw))) return rc;
    QB_CUDA(cudaMemcpyAsync(p->st_kpg2, kpg2, d.ngw * sizeof(double), cudaMemcpyHostToDevice, p->stream));
    kd = p->st_kpg2;
  }
  if (ohost) {
    if ((rc = ensure_buf(&p->st_cp, &p->st_cp_cap, blk))) return rc;
    od = p->st_cp;
  }
  // Host blocks are pipelined by blocks of states: block b+1 is uploaded (copy stream s_in) and block b-1 downloaded
  // (s_out) while block b is computed on the plan's stream; device-resident blocks are one "block".
  const bool pipelined = upload_c || ohost;
  const int SB = pipelined ? hpsi_block_states(nst, nl && qb200_nl_chunks(nl, nst) > 1) : nst;
  const int nblk = (nst + SB - 1) / SB;
  if (pipelined && (rc = plan_copy_streams(p))) return rc;
  if (upload_c) {
    cudaEvent_t ev;
    if ((rc = plan_event(p, 0, &ev))) return rc;
    QB_CUDA(cudaEventRecord(ev, p->stream));
    QB_CUDA(cudaStreamWaitEvent(p->s_in, ev, 0));
    for (int b = 0; b < nblk; b++) {
      const int n0 = b * SB, nb = std::min(SB, nst - n0);
      const size_t off = 2 * (size_t)n0 * ldc;
      QB_CUDA(cudaMemcpyAsync(p->st_c + off, c + off, 2 * (size_t)nb * ldc * sizeof(double), cudaMemcpyHostToDevice, p->s_in));
      if ((rc = plan_event(p, 1 + b, &ev))) return rc;
      QB_CUDA(cudaEventRecord(ev, p->s_in));
    }
  }
  cudaStream_t saved = 0;
  if (nl) saved = qb200_nl_swap_stream(nl, p->stream);
  for (int b = 0; b < nblk; b++) {
    const int n0 = b * SB, nb = std::min(SB, nst - n0);
    const size_t off = 2 * (size_t)n0 * ldc;
    if (upload_c) QB_CUDA(cudaStreamWaitEvent(p->stream, p->evs[1 + b], 0));
    // dwf.c().clear(): with projectors the back-projection WRITES rows [0, ngw) (first term of H psi), so only the
    // padding rows need clearing; without projectors the whole slice is cleared
    const bool nlw = nl && qb200_nl_projectors(nl) > 0;
    if (!nlw) QB_CUDA(cudaMemsetAsync(od + off, 0, 2 * (size_t)nb * ldc * sizeof(double), p->stream));
    else if (ldc > d.ngw)
      QB_CUDA(cudaMemset2DAsync(od + off + 2 * (size_t)d.ngw, (size_t)ldc * 16, 0, (size_t)(ldc - d.ngw) * 16, nb, p->stream));
    if (nl && (rc = qb200_nl_energy_dev(nl, ldc, nb, cd + off, occ + n0, 1, od + off, (b > 0 ? 1 : 0) | (nlw ? 2 : 0)))) {   // nlp->energy(sd, true, dsd, ...)
      qb200_nl_swap_stream(nl, saved);
      return rc;
    }
    if (ev_v && b == 0) QB_CUDA(cudaStreamWaitEvent(p->stream, ev_v, 0));
    if ((rc = qb200_rs_mul_add_dev(p, ldc, nb, cd + off, vd, kd, od + off))) {                       // kinetic + sd.rs_mul_add(...)
      if (nl) qb200_nl_swap_stream(nl, saved);
      return rc;
    }
    if (ohost) {
      cudaEvent_t ev;
      if ((rc = plan_event(p, 1 + nblk + b, &ev))) return rc;
      QB_CUDA(cudaEventRecord(ev, p->stream));
      QB_CUDA(cudaStreamWaitEvent(p->s_out, ev, 0));
      QB_CUDA(cudaMemcpyAsync(hpsi + off, od + off, 2 * (size_t)nb * ldc * sizeof(double), cudaMemcpyDeviceToHost, p->s_out));
    }
  }
  if (nl) qb200_nl_swap_stream(nl, saved);
  double e = 0.0;
  if (nl && enl) QB_CUDA(cudaMemcpyAsync(&e, qb200_nl_enl_dev(nl), sizeof(double), cudaMemcpyDeviceToHost, p->stream));
  if (pipelined || chost || (nl && enl)) QB_CUDA(cudaStreamSynchronize(p->stream));
  if (ohost) QB_CUDA(cudaStreamSynchronize(p->s_out));
  if (upload_c) plan_mark_resident(p, c, ldc, nst);
  if (enl) *enl = e;
  return QB200_OK;
}

// ------------------------------------------------------------------------------------------------ TDDFT propagator glue
// y1 += f1*x ; y2 += f2*x (complex factors), one pass over the block (ComplexMatrix::axpy, ExponentialWavefunctionStepper.cc:124-133)
__global__ void __launch_bounds__(256) k_zaxpy2(size_t n, double2 f1, const double2* __restrict__ x, double2* __restrict__ y1,
                                                double2 f2, double2* __restrict__ y2)
{
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const double2 a = x[i];
    double2 b = y1[i];
    b.x += f1.x * a.x - f1.y * a.y;
    b.y += f1.x * a.y + f1.y * a.x;
    y1[i] = b;
    if (y2) {
      double2 c = y2[i];
      c.x += f2.x * a.x - f2.y * a.y;
      c.y += f2.x * a.y + f2.y * a.x;
      y2[i] = c;
    }
  }
}

extern "C" int qb200_exponential(qb200_plan* p, qb200_nl* nl, int ldc, int nst, double* c, const double* occ, const double* v,
                                 const double* kpg2, int order, double dt1, double dt2, double* c2)
{
  if (!p || !c || !v || nst < 0 || ldc < p->d.ngw || order < 1 || order > 16 || (nl && !occ)) { set_error("qb200_exponential: bad argument"); return QB200_EINVAL; }
  if (nst == 0) return QB200_OK;
  if (p->d.is_real) { set_error("qb200_exponential: the propagator needs complex wavefunctions (the reference requires force_complex_wf, vars/WfDyn.h:82-92)"); return QB200_EINVAL; }
  QB_CUDA(cudaSetDevice(p->device));
  const size_t blk = 2 * (size_t)ldc * nst, nel = (size_t)ldc * nst;
  int rc;
  // device views of c (in/out) and c2 (out)
  const bool chost = !is_device_ptr(c), c2host = c2 && !is_device_ptr(c2);
  double *cd = c, *c2d = c2;
  if ((rc = ensure_buf(&p->ex_a, &p->ex_a_cap, blk)) || (rc = ensure_buf(&p->ex_b, &p->ex_b_cap, blk))) return rc;
  if (chost) {
    p->res_ptr = nullptr;
    if ((rc = ensure_buf(&p->st_c, &p->st_c_cap, blk))) return rc;
    QB_CUDA(cudaMemcpyAsync(p->st_c, c, blk * sizeof(double), cudaMemcpyHostToDevice, p->stream));
    cd = p->st_c;
  }
  if (c2host) {
    if ((rc = ensure_buf(&p->ex_c2, &p->ex_c2_cap, blk))) return rc;
    c2d = p->ex_c2;
  }
  // device copies of v and kpg2 once for all applications of H
  const double *vd = v, *kd = kpg2;
  const size_t N = (size_t)p->d.np0 * p->d.np1 * p->d.np2;
  if (!is_device_ptr(v)) {
    if ((rc = ensure_buf(&p->st_v, &p->st_v_cap, N))) return rc;
    QB_CUDA(cudaMemcpyAsync(p->st_v, v, N * sizeof(double), cudaMemcpyHostToDevice, p->stream));
    vd = p->st_v;
  }
  if (kpg2 && !is_device_ptr(kpg2)) {
    if ((rc = ensure_buf(&p->st_kpg2, &p->st_kpg2_cap, p->d.ngw))) return rc;
    QB_CUDA(cudaMemcpyAsync(p->st_kpg2, kpg2, p->d.ngw * sizeof(double), cudaMemcpyHostToDevice, p->stream));
    kd = p->st_kpg2;
  }
  double2 f1 = make_double2(1.0, 0.0), f2 = make_double2(1.0, 0.0);
  const double* operand = cd;
  double* out = p->ex_b;
  for (int n = 1; n <= order; n++) {
    // factor *= -i dt / n   (ExponentialWavefunctionStepper.cc:100-103)
    f1 = make_double2(f1.y * dt1 / n, -f1.x * dt1 / n);
    f2 = make_double2(f2.y * dt2 / n, -f2.x * dt2 / n);
    if ((rc = qb200_hpsi(p, nl, ldc, nst, operand, occ, vd, kd, out, nullptr))) return rc;      // ef_.energy(wf_, true, dwf, ...)
    if (n == 1 && c2d) QB_CUDA(cudaMemcpyAsync(c2d, cd, blk * sizeof(double), cudaMemcpyDeviceToDevice, p->stream));   // newwf_ = wf_
    k_zaxpy2<<<148 * 8, 256, 0, p->stream>>>(nel, f1, (const double2*)out, (double2*)cd, f2, (double2*)c2d);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "k_zaxpy2 launch", __FILE__, __LINE__);
    p->launches++;
    operand = out;                                                                                // wf_ = dwf
    out = (out == p->ex_b) ? p->ex_a : p->ex_b;
  }
  if (chost) QB_CUDA(cudaMemcpyAsync(c, cd, blk * sizeof(double), cudaMemcpyDeviceToHost, p->stream));
  if (c2host) QB_CUDA(cudaMemcpyAsync(c2, c2d, blk * sizeof(double), cudaMemcpyDeviceToHost, p->stream));
  QB_CUDA(cudaStreamSynchronize(p->stream));
  return QB200_OK;
}

// ------------------------------------------------------------------------------------------------ current density
// CurrentDensity::update_current (CurrentDensity.cc:52-86) with SlaterDet::compute_density(ft, w, complex* rho, sd2)
// (SlaterDet.cc:935-968): per direction d, eta = FT^-1[kpgx_d c] (so that rwf = i*kpgx_d*c transforms to i*eta) and
//   current_d(r) += -Im sum_n fac_n conj(psi_n) (i eta_n) = -sum_n fac_n Re(conj(psi_n) eta_n).
// The reference transforms psi and rwf to the whole grid for every state and direction (6 FFTs per state) and multiplies
// there.  Here the product is taken through the polarisation identity
//   4 s Re(conj(psi) eta) = |psi + s eta|^2 - |psi - s eta|^2 ,        psi +- s eta = FT^-1[(1 +- s kpgx_d) c] ,
// so each direction is TWO runs of the fused density path (z columns -> plane kernel -> |.|^2 accumulation, nothing
// written to the grid per state) on the block scaled per plane wave: the same 6 transforms per state, none of them
// leaving the chip.  s = 1/max|kpgx_d| keeps both weights in [0, 2]; the rounding error is that of two density builds
// (~1e-16 * rho * max|k|, the size of the reference's own rounding).  Real (Gamma) bases: the current of real
// wavefunctions vanishes identically (the reference accumulates rounding noise ~1e-17); nothing is added.
__global__ void __launch_bounds__(1024) k_kmax(const double* __restrict__ kpgx, int ngw, double* __restrict__ kmax)
{
  __shared__ double red[1024];
  const double* k = kpgx + (size_t)blockIdx.x * ngw;
  double m = 0.0;
  for (int i = threadIdx.x; i < ngw; i += 1024) m = fmax(m, fabs(k[i]));
  red[threadIdx.x] = m;
  __syncthreads();
  for (int s = 512; s > 0; s >>= 1) {
    if (threadIdx.x < s) red[threadIdx.x] = fmax(red[threadIdx.x], red[threadIdx.x + s]);
    __syncthreads();
  }
  if (threadIdx.x == 0) kmax[blockIdx.x] = red[0];
}
// out[g, n] = (1 + sign * k[g] / kmax) * c[g, n]; grid (ceil(ngw/256), nst)
__global__ void __launch_bounds__(256) k_scale_k(const double2* __restrict__ c, size_t ldc, int ngw, const double* __restrict__ k,
                                                 const double* __restrict__ kmax, double sign, double2* __restrict__ out)
{
  const int g = blockIdx.x * 256 + threadIdx.x;
  if (g >= ngw) return;
  const double km = *kmax;
  const double w = 1.0 + sign * (km > 0.0 ? k[g] / km : 0.0);
  const double2 a = c[(size_t)blockIdx.y * ldc + g];
  out[(size_t)blockIdx.y * ldc + g] = make_double2(w * a.x, w * a.y);
}
// cur += sign * kmax/4 * rho
__global__ void __launch_bounds__(256) k_cur_acc(double* __restrict__ cur, const double* __restrict__ rho, size_t N,
                                                 const double* __restrict__ kmax, double sign)
{
  const double f = sign * 0.25 * *kmax;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < N; i += (size_t)gridDim.x * blockDim.x) cur[i] += f * rho[i];
}

extern "C" int qb200_compute_current(qb200_plan* p, int ldc, int nst, const double* c, const double* fac, const double* kpgx,
                                     double* cur)
{
  if (!p || !c || !fac || !kpgx || !cur || nst < 0 || ldc < p->d.ngw) { set_error("qb200_compute_current: bad argument"); return QB200_EINVAL; }
  if (nst == 0 || p->d.is_real) return QB200_OK;            // real wavefunctions carry no current
  QB_CUDA(cudaSetDevice(p->device));
  const int ngw = p->d.ngw;
  const size_t N = (size_t)p->d.np0 * p->d.np1 * p->d.np2, blk = 2 * (size_t)ldc * nst;
  int rc;
  // work: the scaled block (ex_a); one grid of density + kmax[3] (+ the device copy of cur and kpgx if they are host arrays)
  const bool chost = !is_device_ptr(c), khost = !is_device_ptr(kpgx), curhost = !is_device_ptr(cur);
  if ((rc = ensure_buf(&p->ex_a, &p->ex_a_cap, blk))) return rc;
  const size_t need = N + 4 + (curhost ? 3 * N : 0) + (khost ? 3 * (size_t)ngw : 0);
  if ((rc = ensure_buf(&p->ex_b, &p->ex_b_cap, need))) return rc;
  double* tmp = p->ex_b;
  double* kmax = tmp + N;
  double* curd = curhost ? tmp + N + 4 : cur;
  const double* kd = kpgx;
  if (khost) {
    double* kbuf = tmp + N + 4 + (curhost ? 3 * N : 0);
    QB_CUDA(cudaMemcpyAsync(kbuf, kpgx, 3 * (size_t)ngw * sizeof(double), cudaMemcpyHostToDevice, p->stream));
    kd = kbuf;
  }
  if (curhost) QB_CUDA(cudaMemcpyAsync(curd, cur, 3 * N * sizeof(double), cudaMemcpyHostToDevice, p->stream));
  const double* cd = c;
  if (chost) {
    p->res_ptr = nullptr;
    if ((rc = ensure_buf(&p->st_c, &p->st_c_cap, blk))) return rc;
    QB_CUDA(cudaMemcpyAsync(p->st_c, c, blk * sizeof(double), cudaMemcpyHostToDevice, p->stream));
    cd = p->st_c;
  }
  // rows >= ngw of the scaled block are padding: cleared once so the block is a valid coefficient block
  if (ldc > ngw) QB_CUDA(cudaMemsetAsync(p->ex_a, 0, blk * sizeof(double), p->stream));
  k_kmax<<<3, 1024, 0, p->stream>>>(kd, ngw, kmax);
  p->launches++;
  const dim3 gs((ngw + 255) / 256, nst);
  for (int d = 0; d < 3; d++)
    for (int pass = 0; pass < 2; pass++) {
      const double sign = pass == 0 ? 1.0 : -1.0;             // |psi + s eta|^2 enters with -, |psi - s eta|^2 with +
      k_scale_k<<<gs, 256, 0, p->stream>>>((const double2*)cd, ldc, ngw, kd + (size_t)d * ngw, kmax + d, sign, (double2*)p->ex_a);
      cudaError_t e = cudaGetLastError();
      if (e != cudaSuccess) return cuda_fail(e, "k_scale_k launch", __FILE__, __LINE__);
      QB_CUDA(cudaMemsetAsync(tmp, 0, N * sizeof(double), p->stream));
      if ((rc = qb200_compute_density(p, ldc, nst, p->ex_a, fac, tmp))) return rc;
      k_cur_acc<<<148 * 4, 256, 0, p->stream>>>(curd + (size_t)d * N, tmp, N, kmax + d, -sign);
      e = cudaGetLastError();
      if (e != cudaSuccess) return cuda_fail(e, "k_cur_acc launch", __FILE__, __LINE__);
      p->launches += 2;
    }
  if (curhost) QB_CUDA(cudaMemcpyAsync(cur, curd, 3 * N * sizeof(double), cudaMemcpyDeviceToHost, p->stream));
  if (curhost || chost) QB_CUDA(cudaStreamSynchronize(p->stream));
  return QB200_OK;
}

// ------------------------------------------------------------------------------------------------ kinetic energy sums (K22)
// EnergyFunctional::energy, "kinetic energy" section (/root/reference/src/qball/EnergyFunctional.cc:1155-1296):
//   psi2sum[ig] = sum_n w[n] |c[ig,n]|^2                  (:1209-1223; w[n] = fac * occ[n], fac = 1 real basis / 0.5 complex)
//   tsum[0]     = sum_ig psi2sum * kpg2                   (:1230)          -> ekin
//   tsum[1..6]  = sum_ig 2 psi2sum * (xx, yy, zz, xy, yz, xz) of k+G       (:1232-1247, compute_stress)
//   tsum[7]     = sum_ig psi2sum * fstress                (:1258)          -> econf
//   tsum[8..13] = sum_ig psi2sum * dfstress * (xx, yy, zz, xy, yz, xz)     (:1260-1273)
// One pass over the block (HBM-bound: 16 ldc nst bytes), states summed in index order per plane wave, then one CTA
// reduces the 14 weighted sums in a fixed order: deterministic.  The k-point weight, the division by weightsum and the
// dsum over ranks (:1281-1294) stay with the caller (qb200_allreduce_scalars).
__global__ void __launch_bounds__(256) k_psi2sum(const double2* __restrict__ c, size_t ldc, int ngw, int nst, const double* __restrict__ w,
                                                 double* __restrict__ psi2sum)
{
  const int ig = blockIdx.x * 256 + threadIdx.x;
  if (ig >= ngw) return;
  double s = 0.0;
  const double2* p = c + ig;
#pragma unroll 4
  for (int n = 0; n < nst; n++) {
    const double2 a = p[(size_t)n * ldc];
    s += w[n] * (a.x * a.x + a.y * a.y);
  }
  psi2sum[ig] = s;
}

__global__ void __launch_bounds__(1024) k_ekin_sums(const double* __restrict__ psi2sum, int ngw, const double* __restrict__ kpg2,
                                                    const double* __restrict__ kpgx, const double* __restrict__ fstress,
                                                    const double* __restrict__ dfstress, double* __restrict__ tsum)
{
  __shared__ double red[1024];
  double t[14];
#pragma unroll
  for (int k = 0; k < 14; k++) t[k] = 0.0;
  for (int ig = threadIdx.x; ig < ngw; ig += 1024) {
    const double p2 = psi2sum[ig];
    t[0] += p2 * kpg2[ig];
    double xx = 0, yy = 0, zz = 0, xy = 0, yz = 0, xz = 0;
    if (kpgx) {
      const double x = kpgx[ig], y = kpgx[ngw + ig], z = kpgx[2 * (size_t)ngw + ig];
      xx = x * x; yy = y * y; zz = z * z; xy = x * y; yz = y * z; xz = x * z;
      const double f = 2.0 * p2;
      t[1] += f * xx; t[2] += f * yy; t[3] += f * zz; t[4] += f * xy; t[5] += f * yz; t[6] += f * xz;
    }
    if (fstress) t[7] += p2 * fstress[ig];
    if (kpgx && dfstress) {
      const double f = p2 * dfstress[ig];
      t[8] += f * xx; t[9] += f * yy; t[10] += f * zz; t[11] += f * xy; t[12] += f * yz; t[13] += f * xz;
    }
  }
  for (int k = 0; k < 14; k++) {
    red[threadIdx.x] = t[k];
    __syncthreads();
    for (int s = 512; s > 0; s >>= 1) {
      if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
      __syncthreads();
    }
    if (threadIdx.x == 0) tsum[k] = red[0];
    __syncthreads();
  }
}

extern "C" int qb200_ekin_sums(qb200_plan* p, int ldc, int nst, const double* c, const double* w, const double* kpg2,
                               const double* kpgx, const double* fstress, const double* dfstress, double* psi2sum, double* tsum)
{
  if (!p || !c || !w || !kpg2 || !tsum || nst < 0 || ldc < p->d.ngw) { set_error("qb200_ekin_sums: bad argument"); return QB200_EINVAL; }
  QB_CUDA(cudaSetDevice(p->device));
  const int ngw = p->d.ngw;
  const size_t blk = 2 * (size_t)ldc * nst;
  int rc;
  for (int k = 0; k < 14; k++) tsum[k] = 0.0;
  if (nst == 0) { if (psi2sum && !is_device_ptr(psi2sum)) for (int i = 0; i < ngw; i++) psi2sum[i] = 0.0; return QB200_OK; }
  // device views: c (resident copy reused under an unchanged coefficient tag), the per-plane-wave tables, the weights
  const double* cd = c;
  if (!is_device_ptr(c)) {
    if ((rc = ensure_buf(&p->st_c, &p->st_c_cap, blk))) return rc;
    if (!plan_resident(p, c, ldc, nst)) {
      p->res_ptr = nullptr;
      QB_CUDA(cudaMemcpyAsync(p->st_c, c, blk * sizeof(double), cudaMemcpyHostToDevice, p->stream));
      plan_mark_resident(p, c, ldc, nst);
    }
    cd = p->st_c;
  }
  // work: psi2sum[ngw] | tsum[14] | w[nst] | host tables staged behind (kpg2, kpgx[3], fstress, dfstress as needed)
  const bool hk = !is_device_ptr(kpg2), hx = kpgx && !is_device_ptr(kpgx), hf = fstress && !is_device_ptr(fstress),
             hd = dfstress && !is_device_ptr(dfstress);
  const size_t need = (size_t)ngw + 16 + nst + (size_t)ngw * ((hk ? 1 : 0) + (hx ? 3 : 0) + (hf ? 1 : 0) + (hd ? 1 : 0));
  if ((rc = ensure_buf(&p->st_f, &p->st_f_cap, need))) return rc;
  double* ps = p->st_f;
  double* ts = ps + ngw;
  double* wd = ts + 16;
  double* nxt = wd + nst;
  auto stage = [&](const double* h, size_t n, bool host, const double** dptr) -> int {
    if (!h) { *dptr = nullptr; return QB200_OK; }
    if (!host) { *dptr = h; return QB200_OK; }
    QB_CUDA(cudaMemcpyAsync(nxt, h, n * sizeof(double), cudaMemcpyHostToDevice, p->stream));
    *dptr = nxt; nxt += n;
    return QB200_OK;
  };
  const double *kd, *xd, *fd, *dd;
  if ((rc = stage(kpg2, ngw, hk, &kd)) || (rc = stage(kpgx, 3 * (size_t)ngw, hx, &xd)) || (rc = stage(fstress, ngw, hf, &fd)) ||
      (rc = stage(dfstress, ngw, hd, &dd))) return rc;
  QB_CUDA(cudaMemcpyAsync(wd, w, nst * sizeof(double), cudaMemcpyDefault, p->stream));
  k_psi2sum<<<(ngw + 255) / 256, 256, 0, p->stream>>>((const double2*)cd, (size_t)ldc, ngw, nst, wd, ps);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "k_psi2sum launch", __FILE__, __LINE__);
  k_ekin_sums<<<1, 1024, 0, p->stream>>>(ps, ngw, kd, xd, fd, dd, ts);
  e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "k_ekin_sums launch", __FILE__, __LINE__);
  p->launches += 2;
  QB_CUDA(cudaMemcpyAsync(tsum, ts, 14 * sizeof(double), cudaMemcpyDefault, p->stream));
  if (psi2sum) QB_CUDA(cudaMemcpyAsync(psi2sum, ps, ngw * sizeof(double), cudaMemcpyDefault, p->stream));
  QB_CUDA(cudaStreamSynchronize(p->stream));
  return QB200_OK;
}
