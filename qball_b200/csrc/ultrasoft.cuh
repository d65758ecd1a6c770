// qball_b200/csrc/ultrasoft.cuh -- included by nonlocal.cu (after the ultrasoft beta.psi entry points).
//
// SURVEY section 8 row f4, the rest of the ultrasoft path:
//   * the ultrasoft branch of NonLocalPotential::energy (src/qball/NonLocalPotential.cc:1554-1752, no forces):
//       D_nm^I = D_nm^0 + sum_G Re( conj(sf_I(G) Q_nm(G)) veff(G) )      (:1607-1636, the !highmem branch; sf of :2724-2731)
//       E_nl   = sum_n occ_n/omega sum_{I,q} mult_q D_q^0 Re( conj(bp_n[I,lm1]) bp_n[I,lm2] )     (:1639-1665)
//       H psi_n += sum_{I,lm} beta^I_lm (1/omega) sum_lm' D^I[lm,lm'] bp_n[I,lm']                   (:1667-1750)
//   * the augmentation charges of ChargeDensity::update_density (src/qball/ChargeDensity.cc:312-465):
//       summat[I,q] = sum_n (weight occ_n/omega) mult_q conj(bp_n[I,lm1]) bp_n[I,lm2]              (:352-368)
//       rhogus(G)   = sum_{I,q} Q_q(G) summat[I,q] exp(-i G.tau_I)/omega                            (:397-428, sfactloc_ :800-820)
//       rho(r)     += Re FT^-1[rhogus]                                                              (:437-456)
// Q_nm(G) on the density basis (Species::calc_qnmg, an input like twnl and betag), D^0, the (lm1, lm2) pairs and veff(G)
// (EnergyFunctional.cc:924-927) come from the caller.  Every sum runs in a fixed order (deterministic).

namespace qb200 {

// part[((ia*nq + q)*gsplit + gs)] = sum over the G of split gs of Re( conj(sf_ia(G) Q_q(G)) veff(G) );
// grid (gsplit, ceil(nq/8), na), 128 threads; the sincos of an (atom, G) is shared by 8 pairs
__global__ void __launch_bounds__(128) k_us_dmat(UsSpeciesDev U, const double* __restrict__ tau, const double* __restrict__ vkpgx, int ngv,
                                                 const double2* __restrict__ veff, int gper, double* __restrict__ part)
{
  __shared__ double red[128];
  const int gs = blockIdx.x, q0 = blockIdx.y * 8, ia = blockIdx.z;
  const double tx = tau[3 * ia], ty = tau[3 * ia + 1], tz = tau[3 * ia + 2];
  const int g0 = gs * gper, g1 = min(ngv, g0 + gper);
  double acc[8];
#pragma unroll
  for (int j = 0; j < 8; j++) acc[j] = 0.0;
  for (int ig = g0 + threadIdx.x; ig < g1; ig += 128) {
    double s, c;
    sincos(tx * vkpgx[ig] + ty * vkpgx[ngv + ig] + tz * vkpgx[2 * (size_t)ngv + ig], &s, &c);
    const double2 v = veff[ig];
    const double ur = c * v.x - s * v.y, ui = c * v.y + s * v.x;        // conj(sf) veff, sf = cos - i sin
#pragma unroll
    for (int j = 0; j < 8; j++) {
      if (q0 + j < U.nq) {
        const double2 qv = U.qnm[(size_t)(q0 + j) * ngv + ig];
        acc[j] += qv.x * ur + qv.y * ui;                                  // Re( conj(Q) u )
      }
    }
  }
  for (int j = 0; j < 8; j++) {
    if (q0 + j >= U.nq) break;
    red[threadIdx.x] = acc[j];
    __syncthreads();
    for (int s = 64; s > 0; s >>= 1) {
      if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
      __syncthreads();
    }
    if (threadIdx.x == 0) part[((size_t)ia * U.nq + q0 + j) * gridDim.x + gs] = red[0];
    __syncthreads();
  }
}

// D[ia][lm1][lm2] = D[ia][lm2][lm1] = dzero_q (+ sum of the splits of k_us_dmat when part != null); one thread per (ia, q)
__global__ void __launch_bounds__(128) k_us_dfull(UsSpeciesDev U, int na, int npr, const double* __restrict__ part, int gsplit,
                                                  double* __restrict__ D)
{
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= na * U.nq) return;
  const int ia = idx / U.nq, q = idx % U.nq;
  double d = 0.0;
  if (part) for (int gs = 0; gs < gsplit; gs++) d += part[(size_t)idx * gsplit + gs];
  d += U.dzero[q];
  const int a = U.lm1[q], b = U.lm2[q];
  double* Da = D + (size_t)ia * npr * npr;
  Da[a * npr + b] = d;
  Da[b * npr + a] = d;
}

// eblk[n] = occ[n] * sum_p Re( conj(bp[n][p]) f[n][p] ); one block per state
__global__ void __launch_bounds__(128) k_us_enl(const double2* __restrict__ bp, const double2* __restrict__ f, int Mtot,
                                                const double* __restrict__ occ, double* __restrict__ eblk)
{
  __shared__ double red[128];
  const int n = blockIdx.x;
  double e = 0.0;
  for (int p = threadIdx.x; p < Mtot; p += 128) {
    const double2 a = bp[(size_t)n * Mtot + p], b = f[(size_t)n * Mtot + p];
    e += a.x * b.x + a.y * b.y;
  }
  red[threadIdx.x] = e;
  __syncthreads();
  for (int s = 64; s > 0; s >>= 1) {
    if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) eblk[n] = occ[n] * red[0];
}

// summat[ia*nq + q] = mult_q sum_n fac[n] conj(bp[n][p0 + ia*npr + lm1]) bp[n][p0 + ia*npr + lm2]; one block per (ia, q)
__global__ void __launch_bounds__(128) k_us_summat(UsSpeciesDev U, int npr, int p0, const double2* __restrict__ bp, int Mtot, int nst,
                                                   const double* __restrict__ fac, double2* __restrict__ summat)
{
  __shared__ double rr[128], ri[128];
  const int q = blockIdx.x, ia = blockIdx.y;
  const int a = p0 + ia * npr + U.lm1[q], b = p0 + ia * npr + U.lm2[q];
  double sr = 0.0, si = 0.0;
  for (int n = threadIdx.x; n < nst; n += 128) {
    const double2 x = bp[(size_t)n * Mtot + a], y = bp[(size_t)n * Mtot + b];
    const double w = fac[n];
    sr += w * (x.x * y.x + x.y * y.y);                                    // conj(x) y
    si += w * (x.x * y.y - x.y * y.x);
  }
  rr[threadIdx.x] = sr; ri[threadIdx.x] = si;
  __syncthreads();
  for (int s = 64; s > 0; s >>= 1) {
    if (threadIdx.x < s) { rr[threadIdx.x] += rr[threadIdx.x + s]; ri[threadIdx.x] += ri[threadIdx.x + s]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const double mult = U.lm1[q] == U.lm2[q] ? 1.0 : 2.0;
    summat[(size_t)ia * U.nq + q] = make_double2(mult * rr[0], mult * ri[0]);
  }
}

// rhog[ig] += omega_inv sum_ia exp(-i G.tau_ia) sum_q Q_q(G) summat[ia][q]; one thread per G
__global__ void __launch_bounds__(128) k_us_rhog(UsSpeciesDev U, int na, const double* __restrict__ tau, const double* __restrict__ vkpgx,
                                                 int ngv, const double2* __restrict__ summat, double omega_inv, double2* __restrict__ rhog)
{
  const int ig = blockIdx.x * blockDim.x + threadIdx.x;
  if (ig >= ngv) return;
  const double gx = vkpgx[ig], gy = vkpgx[ngv + ig], gz = vkpgx[2 * (size_t)ngv + ig];
  double ar = 0.0, ai = 0.0;
  for (int ia = 0; ia < na; ia++) {
    double tr = 0.0, ti = 0.0;
    const double2* sm = summat + (size_t)ia * U.nq;
    for (int q = 0; q < U.nq; q++) {
      const double2 qv = U.qnm[(size_t)q * ngv + ig], s = sm[q];
      tr += qv.x * s.x - qv.y * s.y;
      ti += qv.x * s.y + qv.y * s.x;
    }
    double sn, cs;
    sincos(tau[3 * ia] * gx + tau[3 * ia + 1] * gy + tau[3 * ia + 2] * gz, &sn, &cs);
    ar += tr * cs + ti * sn;                                              // t * (cos - i sin)
    ai += ti * cs - tr * sn;
  }
  double2 r = rhog[ig];
  r.x += omega_inv * ar; r.y += omega_inv * ai;
  rhog[ig] = r;
}

// rho[i] += Re f[i]; blk[b] = sum of Re f over the block's points (fixed order)
__global__ void __launch_bounds__(256) k_us_add_real(const double2* __restrict__ f, size_t N, double* __restrict__ rho, double* __restrict__ blk)
{
  __shared__ double red[256];
  double s = 0.0;
  const size_t per = (N + gridDim.x - 1) / gridDim.x, i0 = blockIdx.x * per, i1 = i0 + per < N ? i0 + per : N;
  for (size_t i = i0 + threadIdx.x; i < i1; i += 256) { const double v = f[i].x; rho[i] += v; s += v; }
  red[threadIdx.x] = s;
  __syncthreads();
  for (int k = 128; k > 0; k >>= 1) {
    if (threadIdx.x < k) red[threadIdx.x] += red[threadIdx.x + k];
    __syncthreads();
  }
  if (threadIdx.x == 0) blk[blockIdx.x] = red[0];
}

}  // namespace qb200

using namespace qb200;

extern "C" int qb200_nl_us_set_density_basis(qb200_nl* nl, int ngv, const double* vkpgx)
{
  if (!nl || ngv < 1 || !vkpgx) { set_error("qb200_nl_us_set_density_basis: bad argument"); return QB200_EINVAL; }
  QB_CUDA(cudaSetDevice(nl->device));
  if (!nl->us) nl->us = new UsTables();
  UsTables& T = *nl->us;
  if (T.ngv && T.ngv != ngv) for (UsSpeciesDev& u : T.sp) u.nq = 0;       // tables of another basis: must be set again
  T.ngv = ngv;
  void* d = nullptr;
  QB_CUDA(cudaMalloc(&d, 3 * (size_t)ngv * sizeof(double)));
  nl->owned.push_back(d);
  QB_CUDA(cudaMemcpy(d, vkpgx, 3 * (size_t)ngv * sizeof(double), cudaMemcpyDefault));
  T.vkpgx = (const double*)d;
  return QB200_OK;
}

extern "C" int qb200_nl_us_set_species(qb200_nl* nl, int is, int nq, const int* lm1, const int* lm2, const double* dzero, const double* qnmg)
{
  if (!nl || is < 0 || is >= (int)nl->sp.size() || nq < 1 || !lm1 || !lm2 || !dzero || !qnmg) { set_error("qb200_nl_us_set_species: bad argument"); return QB200_EINVAL; }
  QB_CUDA(cudaSetDevice(nl->device));
  if (!nl->us || !nl->us->ngv) { set_error("qb200_nl_us_set_species: call qb200_nl_us_set_density_basis first"); return QB200_EINVAL; }
  UsTables& T = *nl->us;
  const int npr = nl->sp[is].npr;
  for (int q = 0; q < nq; q++)
    if (lm1[q] < 0 || lm1[q] >= npr || lm2[q] < 0 || lm2[q] >= npr) { set_error("qb200_nl_us_set_species: channel index out of range"); return QB200_EINVAL; }
  if (T.sp.size() < nl->sp.size()) T.sp.resize(nl->sp.size(), UsSpeciesDev{0, nullptr, nullptr, nullptr, nullptr});
  auto up = [&](const void* h, size_t bytes, const void** out) -> int {
    void* d = nullptr;
    QB_CUDA(cudaMalloc(&d, std::max<size_t>(bytes, 16)));
    nl->owned.push_back(d);
    QB_CUDA(cudaMemcpy(d, h, bytes, cudaMemcpyDefault));
    *out = d;
    return QB200_OK;
  };
  UsSpeciesDev u;
  u.nq = nq;
  int rc;
  if ((rc = up(lm1, nq * sizeof(int), (const void**)&u.lm1)) || (rc = up(lm2, nq * sizeof(int), (const void**)&u.lm2)) ||
      (rc = up(dzero, nq * sizeof(double), (const void**)&u.dzero)) || (rc = up(qnmg, (size_t)nq * T.ngv * 16, (const void**)&u.qnm))) return rc;
  T.sp[is] = u;
  return QB200_OK;
}

// the object's tables, complete for every species that has projectors
static int us_get(qb200_nl* nl, const char* who, UsTables* out)
{
  if (!nl->us || !nl->us->ngv) { set_error(std::string(who) + ": no density basis (qb200_nl_us_set_density_basis)"); return QB200_EINVAL; }
  for (size_t is = 0; is < nl->sp.size(); is++)
    if (nl->sp[is].M > 0 && (is >= nl->us->sp.size() || nl->us->sp[is].nq == 0)) {
      set_error(std::string(who) + ": augmentation tables missing for a species (qb200_nl_us_set_species)"); return QB200_EINVAL;
    }
  *out = *nl->us;
  return QB200_OK;
}

// per-projector description for k_us_couple: block start, npr, offset of the coupling matrix, channel; per_atom: one matrix per atom
static std::vector<int> us_meta(const qb200_nl* nl, bool per_atom, size_t* total)
{
  std::vector<int> meta(4 * (size_t)nl->Mtot);
  size_t qo = 0;
  for (const NlSpecies& S : nl->sp) {
    for (int ia = 0; ia < S.na; ia++)
      for (int lm = 0; lm < S.npr; lm++) {
        const size_t p = (size_t)S.poff + (size_t)ia * S.npr + lm;
        meta[4 * p] = S.poff + ia * S.npr; meta[4 * p + 1] = S.npr;
        meta[4 * p + 2] = (int)(qo + (per_atom ? (size_t)ia * S.npr * S.npr : 0)); meta[4 * p + 3] = lm;
      }
    qo += (size_t)S.npr * S.npr * (per_atom ? S.na : 1);
  }
  *total = qo;
  return meta;
}

struct UsScratch {                       // device scratch of one call, freed on every exit path
  std::vector<void*> p;
  ~UsScratch() { for (void* q : p) cudaFree(q); }
  int get(size_t bytes, void** out) { void* d = nullptr; QB_CUDA(cudaMalloc(&d, std::max<size_t>(bytes, 16))); p.push_back(d); *out = d; return QB200_OK; }
};

extern "C" int qb200_nl_us_energy(qb200_nl* nl, int ldc, int nst, const double* c, const double* occ, const double* veff, int compute_hpsi,
                                  double* cp, double* enl)
{
  if (!nl || !c || !occ || !enl || nst < 0 || ldc < nl->ngw || (compute_hpsi && (!cp || !veff))) { set_error("qb200_nl_us_energy: bad argument"); return QB200_EINVAL; }
  int rc;
  if ((rc = nl_us_check(nl, "qb200_nl_us_energy"))) return rc;
  *enl = 0.0;
  if (nst == 0 || nl->Mtot == 0) return QB200_OK;
  UsTables T;
  if ((rc = us_get(nl, "qb200_nl_us_energy", &T))) return rc;
  QB_CUDA(cudaSetDevice(nl->device));
  const int Mtot = nl->Mtot;
  const size_t blk = 2 * (size_t)ldc * nst, nbp = 2 * (size_t)nst * Mtot;
  const double* cd = c;
  if (!is_device_ptr(c)) {
    if ((rc = nl_ensure(&nl->st_c, &nl->st_c_cap, blk))) return rc;
    QB_CUDA(cudaMemcpyAsync(nl->st_c, c, blk * sizeof(double), cudaMemcpyHostToDevice, nl->stream));
    cd = nl->st_c;
  }
  double* cpd = cp;
  if (compute_hpsi && !is_device_ptr(cp)) {
    if ((rc = nl_ensure(&nl->st_cp, &nl->st_cp_cap, blk))) return rc;
    QB_CUDA(cudaMemcpyAsync(nl->st_cp, cp, blk * sizeof(double), cudaMemcpyHostToDevice, nl->stream));
    cpd = nl->st_cp;
  }
  UsScratch S;
  void *bpd, *fd, *md, *Dd, *occd, *ed, *veffd = nullptr, *partd = nullptr;
  size_t nD0, nD;
  const std::vector<int> meta0 = us_meta(nl, false, &nD0), meta1 = us_meta(nl, true, &nD);
  if ((rc = S.get(nbp * 8, &bpd)) || (rc = S.get(nbp * 8, &fd)) || (rc = S.get(meta0.size() * sizeof(int), &md)) ||
      (rc = S.get(std::max(nD0, nD) * 8, &Dd)) || (rc = S.get((size_t)nst * 8, &occd)) || (rc = S.get((size_t)(nst + 1) * 8, &ed))) return rc;
  QB_CUDA(cudaMemcpyAsync(occd, occ, (size_t)nst * 8, cudaMemcpyDefault, nl->stream));
  if ((rc = nl_us_project(nl, ldc, nst, cd, (double*)bpd))) return rc;             // calc_betapsi (:1556)
  const size_t total = (size_t)nst * Mtot;
  // E_nl: coupling with D^0 alone (:1639-1665)
  QB_CUDA(cudaMemcpyAsync(md, meta0.data(), meta0.size() * sizeof(int), cudaMemcpyHostToDevice, nl->stream));
  QB_CUDA(cudaMemsetAsync(Dd, 0, nD0 * 8, nl->stream));
  {
    size_t qo = 0;
    for (size_t is = 0; is < nl->sp.size(); is++) {
      const NlSpecies& N = nl->sp[is];
      if (N.M <= 0) continue;
      k_us_dfull<<<(T.sp[is].nq + 127) / 128, 128, 0, nl->stream>>>(T.sp[is], 1, N.npr, nullptr, 0, (double*)Dd + qo);
      NL_LAUNCH_CHECK(nl);
      qo += (size_t)N.npr * N.npr;
    }
  }
  k_us_couple<<<(unsigned)((total + 255) / 256), 256, 0, nl->stream>>>((const double2*)bpd, Mtot, nst, (const int*)md, (const double*)Dd, 1.0 / nl->omega, (double2*)fd);
  NL_LAUNCH_CHECK(nl);
  k_us_enl<<<nst, 128, 0, nl->stream>>>((const double2*)bpd, (const double2*)fd, Mtot, (const double*)occd, (double*)ed);
  NL_LAUNCH_CHECK(nl);
  QB_CUDA(cudaMemsetAsync((double*)ed + nst, 0, sizeof(double), nl->stream));
  k_sum_blocks<<<1, 256, 0, nl->stream>>>((const double*)ed, nst, (double*)ed + nst);
  NL_LAUNCH_CHECK(nl);
  QB_CUDA(cudaMemcpyAsync(enl, (double*)ed + nst, sizeof(double), cudaMemcpyDeviceToHost, nl->stream));
  if (compute_hpsi) {
    // D^I for every atom (:1582-1636), then A = D bp / omega and the back-projection (:1667-1750)
    if ((rc = S.get((size_t)T.ngv * 16, &veffd))) return rc;
    QB_CUDA(cudaMemcpyAsync(veffd, veff, (size_t)T.ngv * 16, cudaMemcpyDefault, nl->stream));
    const int gsplit = std::max(1, std::min(64, T.ngv / 2048)), gper = (T.ngv + gsplit - 1) / gsplit;
    size_t npart = 0;
    for (size_t is = 0; is < nl->sp.size(); is++) if (nl->sp[is].M > 0) npart = std::max(npart, (size_t)nl->sp[is].na * T.sp[is].nq * gsplit);
    if ((rc = S.get(npart * 8, &partd))) return rc;
    QB_CUDA(cudaMemcpyAsync(md, meta1.data(), meta1.size() * sizeof(int), cudaMemcpyHostToDevice, nl->stream));
    QB_CUDA(cudaMemsetAsync(Dd, 0, nD * 8, nl->stream));
    size_t qo = 0;
    for (size_t is = 0; is < nl->sp.size(); is++) {
      const NlSpecies& N = nl->sp[is];
      if (N.M <= 0) continue;
      const UsSpeciesDev& U = T.sp[is];
      k_us_dmat<<<dim3(gsplit, (U.nq + 7) / 8, N.na), 128, 0, nl->stream>>>(U, N.tau, T.vkpgx, T.ngv, (const double2*)veffd, gper, (double*)partd);
      NL_LAUNCH_CHECK(nl);
      k_us_dfull<<<(N.na * U.nq + 127) / 128, 128, 0, nl->stream>>>(U, N.na, N.npr, (const double*)partd, gsplit, (double*)Dd + qo);
      NL_LAUNCH_CHECK(nl);
      qo += (size_t)N.npr * N.npr * N.na;
    }
    k_us_couple<<<(unsigned)((total + 255) / 256), 256, 0, nl->stream>>>((const double2*)bpd, Mtot, nst, (const int*)md, (const double*)Dd, 1.0 / nl->omega, (double2*)fd);
    NL_LAUNCH_CHECK(nl);
    if ((rc = nl_us_backproject(nl, ldc, nst, (const double*)fd, cpd))) return rc;
    if (cpd != cp) QB_CUDA(cudaMemcpyAsync(cp, cpd, blk * sizeof(double), cudaMemcpyDeviceToHost, nl->stream));
  }
  QB_CUDA(cudaStreamSynchronize(nl->stream));
  return QB200_OK;
}

extern "C" int qb200_nl_us_augment_density(qb200_nl* nl, qb200_plan* pv, int ldc, int nst, const double* c, const double* fac, double* rho,
                                           double* uscharge)
{
  if (!nl || !pv || !c || !fac || !rho || nst < 0 || ldc < nl->ngw) { set_error("qb200_nl_us_augment_density: bad argument"); return QB200_EINVAL; }
  int rc;
  if ((rc = nl_us_check(nl, "qb200_nl_us_augment_density"))) return rc;
  if (uscharge) *uscharge = 0.0;
  if (nst == 0 || nl->Mtot == 0) return QB200_OK;
  UsTables T;
  if ((rc = us_get(nl, "qb200_nl_us_augment_density", &T))) return rc;
  int np[3];
  for (int d = 0; d < 3; d++) np[d] = (int)qb200_plan_query(pv, d);
  if ((int)qb200_plan_query(pv, 5) != T.ngv) { set_error("qb200_nl_us_augment_density: the plan is not the density basis the tables were given on"); return QB200_EINVAL; }
  QB_CUDA(cudaSetDevice(nl->device));
  const int Mtot = nl->Mtot;
  const size_t N = (size_t)np[0] * np[1] * np[2], blk = 2 * (size_t)ldc * nst, nbp = 2 * (size_t)nst * Mtot;
  const double* cd = c;
  if (!is_device_ptr(c)) {
    if ((rc = nl_ensure(&nl->st_c, &nl->st_c_cap, blk))) return rc;
    QB_CUDA(cudaMemcpyAsync(nl->st_c, c, blk * sizeof(double), cudaMemcpyHostToDevice, nl->stream));
    cd = nl->st_c;
  }
  UsScratch S;
  const int nblk = 148 * 2;
  void *bpd, *facd, *smd, *rgd, *fd, *rhod = rho, *bs;
  size_t nsm = 0;
  for (size_t is = 0; is < nl->sp.size(); is++) if (nl->sp[is].M > 0) nsm = std::max(nsm, (size_t)nl->sp[is].na * T.sp[is].nq);
  if ((rc = S.get(nbp * 8, &bpd)) || (rc = S.get((size_t)nst * 8, &facd)) || (rc = S.get(nsm * 16, &smd)) || (rc = S.get((size_t)T.ngv * 16, &rgd)) ||
      (rc = S.get(N * 16, &fd)) || (rc = S.get((size_t)(nblk + 1) * 8, &bs))) return rc;
  const bool rhost = !is_device_ptr(rho);
  if (rhost) {
    if ((rc = S.get(N * 8, &rhod))) return rc;
    QB_CUDA(cudaMemcpyAsync(rhod, rho, N * 8, cudaMemcpyHostToDevice, nl->stream));
  }
  QB_CUDA(cudaMemcpyAsync(facd, fac, (size_t)nst * 8, cudaMemcpyDefault, nl->stream));
  if ((rc = nl_us_project(nl, ldc, nst, cd, (double*)bpd))) return rc;             // sdp->calc_betapsi() (ChargeDensity.cc:333)
  QB_CUDA(cudaMemsetAsync(rgd, 0, (size_t)T.ngv * 16, nl->stream));
  for (size_t is = 0; is < nl->sp.size(); is++) {
    const NlSpecies& Ns = nl->sp[is];
    if (Ns.M <= 0) continue;
    const UsSpeciesDev& U = T.sp[is];
    k_us_summat<<<dim3(U.nq, Ns.na), 128, 0, nl->stream>>>(U, Ns.npr, Ns.poff, (const double2*)bpd, Mtot, nst, (const double*)facd, (double2*)smd);
    NL_LAUNCH_CHECK(nl);
    k_us_rhog<<<(T.ngv + 127) / 128, 128, 0, nl->stream>>>(U, Ns.na, Ns.tau, T.vkpgx, T.ngv, (const double2*)smd, 1.0 / nl->omega, (double2*)rgd);
    NL_LAUNCH_CHECK(nl);
  }
  QB_CUDA(cudaStreamSynchronize(nl->stream));
  if ((rc = qb200_fft_backward(pv, (const double*)rgd, (double*)fd))) return rc;    // vft_->backward(rhogus, rhotmp) (:437); synchronises pv's stream
  k_us_add_real<<<nblk, 256, 0, nl->stream>>>((const double2*)fd, N, (double*)rhod, (double*)bs);
  NL_LAUNCH_CHECK(nl);
  QB_CUDA(cudaMemsetAsync((double*)bs + nblk, 0, sizeof(double), nl->stream));
  k_sum_blocks<<<1, 256, 0, nl->stream>>>((const double*)bs, nblk, (double*)bs + nblk);
  NL_LAUNCH_CHECK(nl);
  double sum = 0.0;
  QB_CUDA(cudaMemcpyAsync(&sum, (double*)bs + nblk, sizeof(double), cudaMemcpyDeviceToHost, nl->stream));
  if (rhost) QB_CUDA(cudaMemcpyAsync(rho, rhod, N * 8, cudaMemcpyDeviceToHost, nl->stream));
  QB_CUDA(cudaStreamSynchronize(nl->stream));
  if (uscharge) *uscharge = sum * nl->omega / (double)N;                            // (:441-444)
  return QB200_OK;
}
