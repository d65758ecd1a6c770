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
  if (!is_device_ptr(v)) {
    if ((rc = ensure_buf(&p->st_v, &p->st_v_cap, N))) return rc;
    QB_CUDA(cudaMemcpyAsync(p->st_v, v, N * sizeof(double), cudaMemcpyHostToDevice, p->stream));
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
