// qball_b200/csrc/hpsi.cu -- the H psi column block of EnergyFunctional::energy(compute_hpsi = true)
// (/root/reference/src/qball/EnergyFunctional.cc:1142-1153 clear, :1500 nonlocal, :1675-1690 kinetic, :1695 local),
// in the reference's order, on one stream, with host blocks staged in and out when host pointers are given.
#include "qb200_internal.h"
#include <algorithm>

int qb200_rs_mul_add_dev(qb200_plan* p, int ldc, int nst, const double* c, const double* v, const double* kpg2, double* cp);
int qb200_nl_energy_dev(qb200_nl* nl, int ldc, int nst, const double* c, const double* occ_host, int compute_hpsi, double* cp);
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
  if (!is_device_ptr(c)) {
    if ((rc = ensure_buf(&p->st_c, &p->st_c_cap, blk))) return rc;
    QB_CUDA(cudaMemcpyAsync(p->st_c, c, blk * sizeof(double), cudaMemcpyHostToDevice, p->stream));
    cd = p->st_c;
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
  if (!is_device_ptr(hpsi)) {
    if ((rc = ensure_buf(&p->st_cp, &p->st_cp_cap, blk))) return rc;
    od = p->st_cp;
  }
  QB_CUDA(cudaMemsetAsync(od, 0, blk * sizeof(double), p->stream));           // dwf.c().clear()
  cudaStream_t saved = 0;
  if (nl) {
    saved = qb200_nl_swap_stream(nl, p->stream);
    rc = qb200_nl_energy_dev(nl, ldc, nst, cd, occ, 1, od);                    // nlp->energy(sd, true, dsd, ...)
    qb200_nl_swap_stream(nl, saved);
    if (rc) return rc;
  }
  if ((rc = qb200_rs_mul_add_dev(p, ldc, nst, cd, vd, kd, od))) return rc;     // kinetic + sd.rs_mul_add(...)
  if (od != hpsi) QB_CUDA(cudaMemcpyAsync(hpsi, od, blk * sizeof(double), cudaMemcpyDeviceToHost, p->stream));
  double e = 0.0;
  if (nl && enl) QB_CUDA(cudaMemcpyAsync(&e, qb200_nl_enl_dev(nl), sizeof(double), cudaMemcpyDeviceToHost, p->stream));
  if (od != hpsi || (nl && enl)) QB_CUDA(cudaStreamSynchronize(p->stream));
  if (enl) *enl = e;
  return QB200_OK;
}
