// qball_b200/csrc/transform_kernels.cuh
// The transform kernels of the H psi / density path.  One "unit" = one complex FFT = one state, or one PAIR of real
// states at the Gamma point (FourierTransform.cc:555-581, 1684-1752).
//
//   k_zcol_bwd   sphere scatter (vector_to_zvec / doublevector_to_zvec, FourierTransform.cc:1624-1720) + z-FFT(+1);
//                writes the column-form intermediate TRANSPOSED, zt[unit][z][iv], so that the xy stage reads each
//                plane's nvec values contiguously.  Zero columns of the grid are never touched.
//   k_plane2<OP> (plane-fused path: one xy-plane in shared memory) lives in plane_kernels.cuh / plane.cu.
//   k_xrows_*, k_ycols<OP>  (split path for planes larger than shared memory, e.g. Au 252x252): the same three
//                phases as separate kernels with a compact kept-rows intermediate w[unit][z][jr][x].
//   k_zcol_fwd   z-FFT(-1), 1/N scale (FourierTransform.cc:1338-1342), sphere gather (zvec_to_vector /
//                zvec_to_doublevector, :1666-1752), fused with cp += (daxpy/zaxpy, SlaterDet.cc:1005-1036) and the
//                kinetic term 0.5|k+G|^2 c (EnergyFunctional.cc:1675-1690).
#pragma once
#include "qb200_internal.h"

namespace qb200 {

__device__ __forceinline__ int colfirst(const DevPlan& P, int r) { return P.is_real ? (r == 0 ? 0 : 2 * r - 1) : r; }

__device__ __forceinline__ void load_tw(cplx* dst, const cplx* __restrict__ src, int n)
{
  for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = src[i];
}

// ------------------------------------------------------------------------------------------------ z columns, backward
// grid (ceil(nrods/rb), nunits); dynamic smem: np2 + ncolmax*pitch complex
template <int MODE>
__global__ void __launch_bounds__(256, 2) k_zcol_bwd(const __grid_constant__ DevPlan P, const cplx* __restrict__ c, size_t ldc, cplx* __restrict__ zt)
{
  extern __shared__ __align__(16) unsigned char smraw[];
  cplx* tw = reinterpret_cast<cplx*>(smraw);
  cplx* lines = tw + P.np2;
  const int np2 = P.np2, pitch = np2 | 1;
  const int unit = blockIdx.y;
  const int r0 = blockIdx.x * P.rb, r1 = min(r0 + P.rb, P.nrods);
  const int col0 = colfirst(P, r0), ncol = colfirst(P, r1) - col0;
  const cplx* c1 = c + (MODE == MODE_PAIR ? 2 * (size_t)unit : (size_t)unit) * ldc;
  const cplx* c2 = c1 + ldc;
  load_tw(tw, P.tw2, np2);
  for (int i = threadIdx.x; i < ncol * pitch; i += blockDim.x) lines[i] = make_double2(0.0, 0.0);
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  for (int r = r0 + warp; r < r1; r += nwarps) {
    const int first = P.rod_first[r], size = P.rod_size[r], lmin = P.rod_lmin[r];
    const int lp = colfirst(P, r) - col0;
    const int lm = (r == 0) ? lp : lp + 1;
    for (int i = lane; i < size; i += 32) {
      const int l = lmin + i;
      const int izp = l < 0 ? l + np2 : l;
      const int izm = l > 0 ? np2 - l : -l;
      const cplx a = c1[first + i];
      cplx p, m;
      if (MODE == MODE_PAIR) {
        const cplx b = c2[first + i];
        p = make_double2(a.x - b.y, a.y + b.x);
        m = make_double2(a.x + b.y, b.x - a.y);
      } else {
        p = a;
        m = make_double2(a.x, -a.y);
      }
      lines[lp * pitch + izp] = p;
      if (P.is_real) lines[lm * pitch + izm] = m;   // same thread: the conjugate write wins at G=0, as in the reference
    }
  }
  __syncthreads();
  LineMap lmz = { pitch, ncol, 0 };
  fft_lines<+1>(lines, ncol, lmz, 1, P.f2, tw);
  cplx* out = zt + (size_t)unit * np2 * P.nvec + col0;
  for (int e = threadIdx.x; e < ncol * np2; e += blockDim.x) {
    const int lc = e % ncol, z = e / ncol;
    out[(size_t)z * P.nvec + lc] = lines[lc * pitch + z];
  }
}

// ------------------------------------------------------------------------------------------------ z columns, forward
// out1/out2 (+ unit offsets) receive the coefficients; accumulate: out += ; kpg2/cin: add 0.5*kpg2*c
template <int MODE>
__global__ void __launch_bounds__(256, 2) k_zcol_fwd(const __grid_constant__ DevPlan P, const cplx* __restrict__ zt, cplx* __restrict__ out, size_t ldc,
                                                  int accumulate, const double* __restrict__ kpg2,
                                                  const cplx* __restrict__ cin, double scale)
{
  extern __shared__ __align__(16) unsigned char smraw[];
  cplx* tw = reinterpret_cast<cplx*>(smraw);
  cplx* lines = tw + P.np2;
  const int np2 = P.np2, pitch = np2 | 1;
  const int unit = blockIdx.y;
  const int r0 = blockIdx.x * P.rb, r1 = min(r0 + P.rb, P.nrods);
  const int col0 = colfirst(P, r0), ncol = colfirst(P, r1) - col0;
  load_tw(tw, P.tw2, np2);
  const cplx* in = zt + (size_t)unit * np2 * P.nvec + col0;
  for (int e = threadIdx.x; e < ncol * np2; e += blockDim.x) {
    const int lc = e % ncol, z = e / ncol;
    lines[lc * pitch + z] = in[(size_t)z * P.nvec + lc];
  }
  __syncthreads();
  LineMap lmz = { pitch, ncol, 0 };
  fft_lines<-1>(lines, ncol, lmz, 1, P.f2, tw);
  const size_t s1 = (MODE == MODE_PAIR ? 2 * (size_t)unit : (size_t)unit) * ldc;
  cplx* o1 = out + s1;
  cplx* o2 = o1 + ldc;
  const cplx* i1 = cin ? cin + s1 : nullptr;
  const cplx* i2 = cin ? i1 + ldc : nullptr;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  for (int r = r0 + warp; r < r1; r += nwarps) {
    const int first = P.rod_first[r], size = P.rod_size[r], lmin = P.rod_lmin[r];
    const int lp = colfirst(P, r) - col0;
    const int lm = (r == 0) ? lp : lp + 1;
    for (int i = lane; i < size; i += 32) {
      const int l = lmin + i;
      const int izp = l < 0 ? l + np2 : l;
      const int ig = first + i;
      const cplx p = lines[lp * pitch + izp];
      cplx v1, v2;
      if (MODE == MODE_PAIR) {
        const int izm = l > 0 ? np2 - l : -l;
        const cplx m = lines[lm * pitch + izm];
        const double hs = 0.5 * scale;
        v1 = make_double2(hs * (p.x + m.x), hs * (p.y - m.y));
        v2 = make_double2(hs * (p.y + m.y), hs * (m.x - p.x));
      } else {
        v1 = make_double2(scale * p.x, scale * p.y);
      }
      if (kpg2) {
        const double h = 0.5 * kpg2[ig];
        const cplx a = i1[ig];
        v1.x += h * a.x; v1.y += h * a.y;
        if (MODE == MODE_PAIR) { const cplx b = i2[ig]; v2.x += h * b.x; v2.y += h * b.y; }
      }
      if (accumulate) {
        const cplx a = o1[ig];
        v1.x += a.x; v1.y += a.y;
        if (MODE == MODE_PAIR) { const cplx b = o2[ig]; v2.x += b.x; v2.y += b.y; }
      }
      o1[ig] = v1;
      if (MODE == MODE_PAIR) o2[ig] = v2;
    }
  }
}

// ------------------------------------------------------------------------------------------------ split path: x rows
__device__ __forceinline__ int keptrow_to_row(const DevPlan& P, int jr) { return jr < P.ksplit ? jr : jr + P.kskip; }

// grid (ceil(nkeep/rowb), np2, nunits); smem: np0 + rowb*pitch0 complex
template <int DIR>
__global__ void __launch_bounds__(256, 2) k_xrows(const __grid_constant__ DevPlan P, cplx* __restrict__ zt, cplx* __restrict__ w, int rowb)
{
  extern __shared__ __align__(16) unsigned char smraw[];
  cplx* tw0 = reinterpret_cast<cplx*>(smraw);
  cplx* rows = tw0 + P.np0;
  const int np0 = P.np0, pitch = P.pitch0;
  const int jr0 = blockIdx.x * rowb, jr1 = min(jr0 + rowb, P.nkeep), nr = jr1 - jr0;
  const int z = blockIdx.y, unit = blockIdx.z;
  load_tw(tw0, P.tw0, np0);
  cplx* ztrow = zt + ((size_t)unit * P.np2 + z) * P.nvec;
  cplx* wz = w + (((size_t)unit * P.np2 + z) * P.nkeep + jr0) * np0;
  const int i0 = P.keeprowstart[jr0], i1 = P.keeprowstart[jr1];
  const LineMap lm = { pitch, nr, 0 };
  if (DIR > 0) {
    for (int i = threadIdx.x; i < nr * pitch; i += blockDim.x) rows[i] = make_double2(0.0, 0.0);
    __syncthreads();
    for (int i = i0 + threadIdx.x; i < i1; i += blockDim.x) {
      const int iv = P.keepcols[i];
      const int hk = P.colhk[iv];
      const int hp = hk % np0, kp = hk / np0;
      const int jr = kp < P.ksplit ? kp : kp - P.kskip;
      rows[(jr - jr0) * pitch + hp] = ztrow[iv];
    }
    __syncthreads();
    fft_lines<+1>(rows, nr, lm, 1, P.f0, tw0);
    for (int e = threadIdx.x; e < nr * np0; e += blockDim.x) wz[e] = rows[(e / np0) * pitch + e % np0];
  } else {
    for (int e = threadIdx.x; e < nr * np0; e += blockDim.x) rows[(e / np0) * pitch + e % np0] = wz[e];
    __syncthreads();
    fft_lines<-1>(rows, nr, lm, 1, P.f0, tw0);
    for (int i = i0 + threadIdx.x; i < i1; i += blockDim.x) {
      const int iv = P.keepcols[i];
      const int hk = P.colhk[iv];
      const int hp = hk % np0, kp = hk / np0;
      const int jr = kp < P.ksplit ? kp : kp - P.kskip;
      ztrow[iv] = rows[(jr - jr0) * pitch + hp];
    }
  }
}

// ------------------------------------------------------------------------------------------------ split path: y columns
// grid (ceil(np0/xb), np2, ngroups); smem: np1 + np1*xb complex
template <int OP>
__global__ void __launch_bounds__(256, 2) k_ycols(const __grid_constant__ DevPlan P, cplx* __restrict__ w, const double* __restrict__ v,
                                               cplx* __restrict__ f, double* __restrict__ rho_part,
                                               const double* __restrict__ fac, int nunits, int units_per_group,
                                               int zero_imag)
{
  extern __shared__ __align__(16) unsigned char smraw[];
  cplx* tw1 = reinterpret_cast<cplx*>(smraw);
  cplx* sm = tw1 + P.np1;
  const int np0 = P.np0, np1 = P.np1, np01 = np0 * np1, xb = P.xb;
  const int x0 = blockIdx.x * xb, nx = min(xb, np0 - x0);
  const int z = blockIdx.y;
  const size_t N = (size_t)np01 * P.np2;
  load_tw(tw1, P.tw1, np1);
  const LineMap cols = { 1, nx, 0 };
  const int u0 = blockIdx.z * units_per_group;
  const int u1 = min(u0 + units_per_group, nunits);
  for (int unit = u0; unit < u1; unit++) {
    double facu = 0.0, facv = 0.0;
    if (OP == OP_DENSITY) { if (!fac_active(P, fac, unit)) continue; facu = fac_first(fac, unit); facv = fac_second(P, fac, unit); }
    cplx* wz = w + ((size_t)unit * P.np2 + z) * P.nkeep * np0 + x0;
    __syncthreads();
    if (OP != OP_FWD) {
      if (P.nkeep < np1)
        for (int i = threadIdx.x; i < P.kskip * xb; i += blockDim.x) sm[P.ksplit * xb + i] = make_double2(0.0, 0.0);
      for (int e = threadIdx.x; e < P.nkeep * nx; e += blockDim.x) {
        const int xl = e % nx, jr = e / nx;
        sm[keptrow_to_row(P, jr) * xb + xl] = wz[(size_t)jr * np0 + xl];
      }
      __syncthreads();
      fft_lines<+1>(sm, nx, cols, xb, P.f1, tw1);
    }
    if (OP == OP_HPSI) {
      const double* vz = v + (size_t)z * np01 + x0;
      for (int e = threadIdx.x; e < np1 * nx; e += blockDim.x) {
        const int xl = e % nx, y = e / nx;
        const double vv = vz[(size_t)y * np0 + xl];
        cplx t = sm[y * xb + xl];
        t.x *= vv;
        t.y = zero_imag ? 0.0 : t.y * vv;
        sm[y * xb + xl] = t;
      }
      __syncthreads();
    } else if (OP == OP_DENSITY) {
      double* rz = rho_part + (size_t)blockIdx.z * N + (size_t)z * np01 + x0;
      for (int e = threadIdx.x; e < np1 * nx; e += blockDim.x) {
        const int xl = e % nx, y = e / nx;
        const cplx t = sm[y * xb + xl];
        rz[(size_t)y * np0 + xl] += facu * t.x * t.x + facv * t.y * t.y;
      }
    } else if (OP == OP_BWD) {
      cplx* fz = f + (size_t)unit * N + (size_t)z * np01 + x0;
      for (int e = threadIdx.x; e < np1 * nx; e += blockDim.x) {
        const int xl = e % nx, y = e / nx;
        fz[(size_t)y * np0 + xl] = sm[y * xb + xl];
      }
    } else if (OP == OP_FWD) {
      const cplx* fz = f + (size_t)unit * N + (size_t)z * np01 + x0;
      for (int e = threadIdx.x; e < np1 * nx; e += blockDim.x) {
        const int xl = e % nx, y = e / nx;
        sm[y * xb + xl] = fz[(size_t)y * np0 + xl];
      }
      __syncthreads();
    }
    if (OP == OP_HPSI || OP == OP_FWD) {
      fft_lines<-1>(sm, nx, cols, xb, P.f1, tw1);
      for (int e = threadIdx.x; e < P.nkeep * nx; e += blockDim.x) {
        const int xl = e % nx, jr = e / nx;
        wz[(size_t)jr * np0 + xl] = sm[keptrow_to_row(P, jr) * xb + xl];
      }
    }
  }
}

// flag[0] |= 1 when some state's G = 0 coefficient (row 0 of its column) has an imaginary part that is not rounding noise
// next to the largest of the column's first rows: |Im c_n(0)| > 1e-13 max_{i < 64} |c_n(i)|_inf.  One thread per state.
__global__ void k_gamma_heads(const cplx* __restrict__ c, size_t ldc, int ngw, int nst, int* __restrict__ flag)
{
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= nst) return;
  const cplx* col = c + (size_t)n * ldc;
  double ref = 0.0;
  for (int i = 0; i < min(64, ngw); i++) ref = fmax(ref, fmax(fabs(col[i].x), fabs(col[i].y)));
  if (fabs(col[0].y) > 1e-13 * ref) atomicOr(flag, 1);
}

// rho[i] += sum_g part[g][i], fixed order
__global__ void k_rho_reduce(double* __restrict__ rho, const double* __restrict__ part, size_t N, int ngroups)
{
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < N; i += (size_t)gridDim.x * blockDim.x) {
    double s = 0.0;
    for (int g = 0; g < ngroups; g++) s += part[(size_t)g * N + i];
    rho[i] += s;
  }
}

// rhotmp[i] = (omega*rho[i], 0) and per-block partial sums of rho (ChargeDensity.cc:520-528); fixed grid -> fixed order
__global__ void __launch_bounds__(256) k_rho_expand(const double* __restrict__ rho, size_t N, double omega, cplx* __restrict__ f,
                                                    double* __restrict__ blocksum)
{
  __shared__ double red[256];
  double s = 0.0;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < N; i += (size_t)gridDim.x * blockDim.x) {
    const double r = rho[i];
    s += r;
    f[i] = make_double2(omega * r, 0.0);
  }
  red[threadIdx.x] = s;
  __syncthreads();
  for (int k = 128; k > 0; k >>= 1) {
    if (threadIdx.x < k) red[threadIdx.x] += red[threadIdx.x + k];
    __syncthreads();
  }
  if (threadIdx.x == 0) blocksum[blockIdx.x] = red[0];
}
__global__ void k_sum_fixed(const double* __restrict__ blocksum, int n, double scale, double* __restrict__ out)
{
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    double s = 0.0;
    for (int i = 0; i < n; i++) s += blocksum[i];
    *out = s * scale;
  }
}

}  // namespace qb200
