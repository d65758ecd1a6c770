// qball_b200/csrc/vhxc.cuh -- v(r) producers on the density basis, included by transform.cu:
//   EnergyFunctional::update_vhxc   /root/reference/src/qball/EnergyFunctional.cc:353-975  (one spin, no ESM / NLCC / enthalpy)
//   XCPotential::update             /root/reference/src/qball/XCPotential.cc:104-460       (LDA and PBE, unpolarized)
//   LDAFunctional::xc_unpolarized   /root/reference/src/functionals/LDAFunctional.cc:96-161
//   PBEFunctional::excpbe / gcor2   /root/reference/src/functionals/PBEFunctional.cc:196-291, 482-492
// rho(r), rho(G) -> v_r(r) = v_xc + FT^-1[ v_ion,local(G) + 4 pi (rho_el(G) + rho_ps(G)) / G^2 ] and E_xc, E_ps, E_Hartree,
// entirely on the density-basis plan, so that between the density build and the next H psi nothing visits the host.
//
// GGA: the reference runs 3 backward transforms for grad rho, then per direction forward / multiply by i G_j / backward
// (10 transforms with the local potential).  Here the three i G_j FT[grad_j rho * vxc2] terms are summed in G space together
// with the local potential and transformed back ONCE (linear: same result to rounding): 3 + 3 + 1 transforms.
#pragma once

namespace qb200 {

__device__ __forceinline__ void xc_lda_unpolarized(double rh, double& ee, double& vv)
{
  // Perdew-Zunger parametrisation of Ceperley-Alder, C and D fixed by continuity at rs = 1 (LDAFunctional.cc:96-161)
  const double c1 = 0.6203504908994001, c3 = -0.610887057711;
  const double A = 0.0311, B = -0.048, b1 = 1.0529, b2 = 0.3334, G = -0.1423;
  const double D = G / (1.0 + b1 + b2) - B;
  const double C = -A - D - G * ((b1 / 2.0 + b2) / ((1.0 + b1 + b2) * (1.0 + b1 + b2)));
  ee = 0.0; vv = 0.0;
  if (rh > 0.0) {
    const double rs = c1 / cbrt(rh);
    const double vx = c3 / rs, ex = 0.75 * vx;
    double ec, vc;
    if (rs < 1.0) {
      const double logrs = log(rs);
      ec = A * logrs + B + C * rs * logrs + D * rs;
      vc = A * logrs + (B - A / 3.0) + (2.0 / 3.0) * C * rs * logrs + ((2.0 * D - C) / 3.0) * rs;
    } else {
      const double sqrtrs = sqrt(rs), den = 1.0 + b1 * sqrtrs + b2 * rs;
      ec = G / den;
      vc = ec * (1.0 + (7.0 / 6.0) * b1 * sqrtrs + (4.0 / 3.0) * b2 * rs) / den;
    }
    ee = ex + ec; vv = vx + vc;
  }
}

__device__ __forceinline__ void pbe_gcor2(double a, double a1, double b1, double b2, double b3, double b4, double rtrs, double& gg, double& ggrs)
{
  const double q0 = -2.0 * a * (1.0 + a1 * rtrs * rtrs);
  const double q1 = 2.0 * a * rtrs * (b1 + rtrs * (b2 + rtrs * (b3 + rtrs * b4)));
  const double q2 = log(1.0 + 1.0 / q1);
  gg = q0 * q2;
  const double q3 = a * (b1 / rtrs + 2.0 * b2 + rtrs * (3.0 * b3 + 4.0 * b4 * rtrs));
  ggrs = -2.0 * a * a1 * q2 - q0 * q3 / (q1 * (1.0 + q1));
}

__device__ __forceinline__ void xc_pbe_unpolarized(double rho, double grad, double& exc, double& vxc1, double& vxc2)
{
  // PBEFunctional::excpbe (PBEFunctional.cc:196-291)
  const double third = 1.0 / 3.0, third4 = 4.0 / 3.0;
  const double ax = -0.7385587663820224058, um = 0.2195149727645171, uk = 0.804, ul = um / uk;
  const double pi32third = 3.09366772628014, alpha = 1.91915829267751, seven_sixth = 7.0 / 6.0, four_over_pi = 1.27323954473516;
  const double gamma = 0.03109069086965489, bet = 0.06672455060314922, delt = bet / gamma;
  exc = 0.0; vxc1 = 0.0; vxc2 = 0.0;
  if (rho < 1.e-18) return;
  const double rh13 = pow(rho, third);
  const double exunif = ax * rh13;
  const double fk = pi32third * rh13;
  const double s = grad / (2.0 * fk * rho);
  const double s2 = s * s, p0 = 1.0 + ul * s2, fxpbe = 1.0 + uk - uk / p0;
  const double ex = exunif * fxpbe;
  const double fs = 2.0 * uk * ul / (p0 * p0);
  const double vx1 = third4 * exunif * (fxpbe - s2 * fs);
  const double vx2 = -exunif * fs / (rho * 4.0 * fk * fk);
  const double rs = alpha / fk;
  const double twoks = 2.0 * sqrt(four_over_pi * fk);
  const double t = grad / (twoks * rho);
  const double rtrs = sqrt(rs);
  double ec, ecrs;
  pbe_gcor2(0.0310907, 0.2137, 7.5957, 3.5876, 1.6382, 0.49294, rtrs, ec, ecrs);
  const double vc = ec - rs * ecrs * third;
  const double pon = -ec / gamma;
  const double b = delt / (exp(pon) - 1.0);
  const double b2 = b * b, t2 = t * t, t4 = t2 * t2;
  const double q4 = 1.0 + b * t2, q5 = q4 + b2 * t4;
  const double h = gamma * log(1.0 + delt * q4 * t2 / q5);
  const double t6 = t4 * t2, rsthrd = rs * third, fac = delt / b + 1.0, bec = b2 * fac / bet;
  const double q8 = q5 * q5 + delt * q4 * q5 * t2, q9 = 1.0 + 2.0 * b * t2;
  const double hb = -bet * b * t6 * (2.0 + b * t2) / q8;
  const double hrs = -rsthrd * hb * bec * ecrs;
  const double ht = 2.0 * bet * q9 / q8;
  const double vc1 = vc + h + hrs - t2 * ht * seven_sixth;
  const double vc2 = -ht / (rho * twoks * twoks);
  exc = ex + ec + h; vxc1 = vx1 + vc1; vxc2 = vx2 + vc2;
}

// fixed-order block reduction of up to NV values per thread into part[blockIdx.x * NV + k]
template <int NV> __device__ __forceinline__ void block_partials(double (&t)[NV], double* __restrict__ part)
{
  __shared__ double red[256];
  for (int k = 0; k < NV; k++) {
    red[threadIdx.x] = t[k];
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
      if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
      __syncthreads();
    }
    if (threadIdx.x == 0) part[(size_t)blockIdx.x * NV + k] = red[0];
    __syncthreads();
  }
}
// out[k] = scale * sum_b part[b*NV + k] in index order (one warp)
__global__ void k_vh_sum(const double* __restrict__ part, int nblk, int nv, double* __restrict__ out)
{
  if (threadIdx.x < nv) {
    double s = 0.0;
    for (int b = 0; b < nblk; b++) s += part[(size_t)b * nv + threadIdx.x];
    out[threadIdx.x] = s;
  }
}

// XC_KIND 0: LDA -- v_r = vxc, partial sum of rho*exc.  1: PBE -- v_r = vxc1, v2 = vxc2 (grad |.| from gr[3][N])
template <int XC_KIND>
__global__ void __launch_bounds__(256) k_vh_xc(const double* __restrict__ rho, const double* __restrict__ gr, size_t N,
                                               double* __restrict__ v_r, double* __restrict__ v2, double* __restrict__ part)
{
  double t[1] = { 0.0 };
  for (size_t i = blockIdx.x * (size_t)256 + threadIdx.x; i < N; i += (size_t)gridDim.x * 256) {
    const double rh = rho[i];
    double e, v1;
    if (XC_KIND == 0) {
      xc_lda_unpolarized(rh, e, v1);
    } else {
      const double gx = gr[i], gy = gr[N + i], gz = gr[2 * N + i];
      double w2;
      xc_pbe_unpolarized(rh, sqrt(gx * gx + gy * gy + gz * gz), e, v1, w2);
      v2[i] = w2;
    }
    v_r[i] = v1;
    t[0] += rh * e;
  }
  block_partials<1>(t, part);
}
// tmp[ig] = i * (gx_j[ig] * scale) * src[ig]
__global__ void __launch_bounds__(256) k_vh_igx(const double2* __restrict__ src, const double* __restrict__ gxj, int ng, double scale,
                                                double2* __restrict__ dst, int accumulate)
{
  const int ig = blockIdx.x * 256 + threadIdx.x;
  if (ig >= ng) return;
  const double g = gxj[ig] * scale;
  const double2 a = src[ig];
  double2 r = make_double2(-g * a.y, g * a.x);
  if (accumulate) { const double2 o = dst[ig]; r.x += o.x; r.y += o.y; }
  dst[ig] = r;
}
__global__ void __launch_bounds__(256) k_vh_take_real(const double2* __restrict__ f, size_t N, double* __restrict__ out)
{
  for (size_t i = blockIdx.x * (size_t)256 + threadIdx.x; i < N; i += (size_t)gridDim.x * 256) out[i] = f[i].x;
}
__global__ void __launch_bounds__(256) k_vh_mul_complex(const double* __restrict__ a, const double* __restrict__ b, size_t N, double2* __restrict__ f)
{
  for (size_t i = blockIdx.x * (size_t)256 + threadIdx.x; i < N; i += (size_t)gridDim.x * 256) f[i] = make_double2(a[i] * b[i], 0.0);
}
__global__ void __launch_bounds__(256) k_vh_add_real(const double2* __restrict__ f, size_t N, double* __restrict__ v_r)
{
  for (size_t i = blockIdx.x * (size_t)256 + threadIdx.x; i < N; i += (size_t)gridDim.x * 256) v_r[i] += f[i].x;
}
// EnergyFunctional.cc:366-379, 447-518: rhoelg = rhog/omega; rhogt = rhoelg + rhopst; vlocal_g = vion + 4 pi rhogt g2i (+ the GGA
// term already in vloc when gga); partial sums: [0] sum Re(conj(rhoelg) vion) weighted (real basis: 2, G = 0 once), [1] sum |rhogt|^2 g2i
__global__ void __launch_bounds__(256) k_vh_local(const double2* __restrict__ rhog, const double2* __restrict__ vion, const double2* __restrict__ rhopst,
                                                  const double* __restrict__ g2i, int ng, double omega_inv, int is_real, int gga,
                                                  double2* __restrict__ vloc, double2* __restrict__ rhogt_out, double* __restrict__ part)
{
  const double fpi = 4.0 * 3.14159265358979323846;
  double t[2] = { 0.0, 0.0 };
  const int ig = blockIdx.x * 256 + threadIdx.x;
  if (ig < ng) {
    const double2 rg = rhog[ig];
    const double2 re = make_double2(omega_inv * rg.x, omega_inv * rg.y);
    const double2 vi = vion[ig], ps = rhopst[ig];
    const double2 rt = make_double2(re.x + ps.x, re.y + ps.y);
    const double gi = g2i[ig];
    double2 vl = make_double2(vi.x + fpi * rt.x * gi, vi.y + fpi * rt.y * gi);
    if (gga) { const double2 o = vloc[ig]; vl.x += o.x; vl.y += o.y; }
    vloc[ig] = vl;
    if (rhogt_out) rhogt_out[ig] = rt;
    const double w = is_real ? (ig == 0 ? 1.0 : 2.0) : 1.0;
    t[0] = w * (re.x * vi.x + re.y * vi.y);
    t[1] = (rt.x * rt.x + rt.y * rt.y) * gi;
  }
  block_partials<2>(t, part);
}

}  // namespace qb200

using namespace qb200;

extern "C" int qb200_update_vhxc(qb200_plan* p, int xc, const double* rhor, const double* rhog, const double* gx, const double* g2i,
                                 const double* vion_local_g, const double* rhopst, double omega, double* v_r, double* rhogt,
                                 double* energies)
{
  if (!p || !rhor || !rhog || !g2i || !vion_local_g || !rhopst || !v_r || !energies || !(omega > 0.0) || (xc != QB200_XC_LDA && xc != QB200_XC_PBE) ||
      (xc == QB200_XC_PBE && !gx)) { set_error("qb200_update_vhxc: bad argument"); return QB200_EINVAL; }
  QB_CUDA(cudaSetDevice(p->device));
  const DevPlan& d = p->d;
  const size_t N = (size_t)d.np0 * d.np1 * d.np2;
  const int ng = d.ngw;
  const bool gga = xc == QB200_XC_PBE;
  const int nblk = 148 * 4, gblk = (ng + 255) / 256;
  // work (doubles): f[2N] | gr[3N] v2[N] (GGA) | tmpg[2ng] vloc[2ng] | part[...] sums[4] | staged host inputs
  auto hostp = [](const double* q) { return q && !is_device_ptr(q); };
  size_t need = 2 * N + (gga ? 4 * N : 0) + 4 * (size_t)ng + (size_t)std::max(nblk, 2 * gblk) + 8;
  need += (hostp(rhor) ? N : 0) + (hostp(rhog) ? 2 * (size_t)ng : 0) + (hostp(gx) ? 3 * (size_t)ng : 0) + (hostp(g2i) ? ng : 0) +
          (hostp(vion_local_g) ? 2 * (size_t)ng : 0) + (hostp(rhopst) ? 2 * (size_t)ng : 0) + (hostp(v_r) ? N : 0) + (hostp(rhogt) ? 2 * (size_t)ng : 0);
  int rc;
  if ((rc = ensure(&p->vh, &p->vh_cap, need))) return rc;
  double* w = p->vh;
  double* f = w; w += 2 * N;
  double *gr = nullptr, *v2 = nullptr;
  if (gga) { gr = w; w += 3 * N; v2 = w; w += N; }
  double* tmpg = w; w += 2 * (size_t)ng;
  double* vloc = w; w += 2 * (size_t)ng;
  double* part = w; w += std::max(nblk, 2 * gblk);
  double* sums = w; w += 8;
  auto in = [&](const double* q, size_t n) -> const double* {
    if (!hostp(q)) return q;
    cudaMemcpyAsync(w, q, n * sizeof(double), cudaMemcpyHostToDevice, p->stream);
    const double* r = w; w += n;
    return r;
  };
  const double* rho_d = in(rhor, N);
  const double* rhog_d = in(rhog, 2 * (size_t)ng);
  const double* gx_d = in(gx, 3 * (size_t)ng);
  const double* g2i_d = in(g2i, ng);
  const double* vion_d = in(vion_local_g, 2 * (size_t)ng);
  const double* ps_d = in(rhopst, 2 * (size_t)ng);
  double* vr_d = v_r;
  if (hostp(v_r)) { vr_d = w; w += N; }
  double* rt_d = rhogt;
  if (hostp(rhogt)) { rt_d = w; w += 2 * (size_t)ng; }
  QB_CUDA(cudaGetLastError());
  const double omega_inv = 1.0 / omega;
  // ---- exchange-correlation (XCPotential::update)
  if (gga) {
    for (int j = 0; j < 3; j++) {                       // grad_j rho = Re FT^-1[ i G_j rho(G) / omega ]      (XCPotential.cc:200-214)
      k_vh_igx<<<gblk, 256, 0, p->stream>>>((const double2*)rhog_d, gx_d + (size_t)j * ng, ng, omega_inv, (double2*)tmpg, 0);
      QB_LAUNCH_CHECK(p);
      if ((rc = fft_backward_impl(p, tmpg, nullptr, f))) return rc;
      k_vh_take_real<<<nblk, 256, 0, p->stream>>>((const double2*)f, N, gr + (size_t)j * N);
      QB_LAUNCH_CHECK(p);
    }
    k_vh_xc<1><<<nblk, 256, 0, p->stream>>>(rho_d, gr, N, vr_d, v2, part);
    QB_LAUNCH_CHECK(p);
  } else {
    k_vh_xc<0><<<nblk, 256, 0, p->stream>>>(rho_d, nullptr, N, vr_d, nullptr, part);
    QB_LAUNCH_CHECK(p);
  }
  k_vh_sum<<<1, 32, 0, p->stream>>>(part, nblk, 1, sums);
  QB_LAUNCH_CHECK(p);
  if (gga) {
    for (int j = 0; j < 3; j++) {                       // vloc(G) (+)= i G_j FT[ grad_j rho * vxc2 ]            (XCPotential.cc:262-285)
      k_vh_mul_complex<<<nblk, 256, 0, p->stream>>>(gr + (size_t)j * N, v2, N, (double2*)f);
      QB_LAUNCH_CHECK(p);
      if ((rc = fft_forward_impl(p, f, tmpg, nullptr))) return rc;
      k_vh_igx<<<gblk, 256, 0, p->stream>>>((const double2*)tmpg, gx_d + (size_t)j * ng, ng, 1.0, (double2*)vloc, j > 0 ? 1 : 0);
      QB_LAUNCH_CHECK(p);
    }
  }
  // ---- local potential, E_ps and E_Hartree (EnergyFunctional.cc:447-518), then ONE backward transform for everything in G space
  k_vh_local<<<gblk, 256, 0, p->stream>>>((const double2*)rhog_d, (const double2*)vion_d, (const double2*)ps_d, g2i_d, ng, omega_inv, d.is_real,
                                           gga ? 1 : 0, (double2*)vloc, (double2*)rt_d, part);
  QB_LAUNCH_CHECK(p);
  k_vh_sum<<<1, 32, 0, p->stream>>>(part, gblk, 2, sums + 1);
  QB_LAUNCH_CHECK(p);
  if ((rc = fft_backward_impl(p, vloc, nullptr, f))) return rc;
  k_vh_add_real<<<nblk, 256, 0, p->stream>>>((const double2*)f, N, vr_d);
  QB_LAUNCH_CHECK(p);
  double h[3];
  QB_CUDA(cudaMemcpyAsync(h, sums, 3 * sizeof(double), cudaMemcpyDeviceToHost, p->stream));
  if (vr_d != v_r) QB_CUDA(cudaMemcpyAsync(v_r, vr_d, N * sizeof(double), cudaMemcpyDeviceToHost, p->stream));
  if (rhogt && rt_d != rhogt) QB_CUDA(cudaMemcpyAsync(rhogt, rt_d, 2 * (size_t)ng * sizeof(double), cudaMemcpyDeviceToHost, p->stream));
  QB_CUDA(cudaStreamSynchronize(p->stream));
  const double fpi = 4.0 * 3.14159265358979323846;
  energies[0] = h[0] * omega / (double)N;                          // exc   (XCPotential.cc:170 / :452)
  energies[1] = h[1] * omega;                                      // eps   (EnergyFunctional.cc:447-465)
  energies[2] = (d.is_real ? 1.0 : 0.5) * omega * fpi * h[2];      // ehart (:497-499)
  return QB200_OK;
}
