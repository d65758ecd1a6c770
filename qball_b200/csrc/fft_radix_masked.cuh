// qball_b200/csrc/fft_radix_masked.cuh
// Butterflies that know at compile time which of their inputs are zero (MASK bit k set <=> x[k] may be non-zero;
// x[k] of a cleared bit is never read).  Used by the shape-specialised plane kernels: the first pass of a pruned
// transform (FourierTransform.cc:202, 772-819: only 2*ntrans0 of np1 rows / |h| <= hmax of np0 columns are non-zero)
// feeds known zeros into its butterflies, and IEEE arithmetic does not let the compiler drop "x + 0.0" on its own.
// Unneeded OUTPUTS need no special code: the compiler removes the dead arithmetic when they are not stored.
#pragma once
#include "fft_radix.cuh"

namespace qb200 {

__device__ __forceinline__ cplx cneg(cplx a) { return make_double2(-a.x, -a.y); }
// a + b / a - b with compile-time-known zero operands (the flags are constants after inlining)
__device__ __forceinline__ cplx zadd(bool za, bool zb, cplx a, cplx b)
{
  if (za && zb) return make_double2(0.0, 0.0);
  if (za) return b;
  if (zb) return a;
  return cadd(a, b);
}
__device__ __forceinline__ cplx zsub(bool za, bool zb, cplx a, cplx b)
{
  if (za && zb) return make_double2(0.0, 0.0);
  if (za) return cneg(b);
  if (zb) return a;
  return csub(a, b);
}

__host__ __device__ constexpr bool mask_bit(unsigned mask, int k) { return ((mask >> k) & 1u) != 0; }

template <int R, int S, unsigned MASK> struct DftM;

template <int S, unsigned MASK> struct DftM<1, S, MASK> { static __device__ __forceinline__ void run(cplx*) {} };

template <int S, unsigned MASK> struct DftM<2, S, MASK> {
  static __device__ __forceinline__ void run(cplx* x)
  {
    constexpr bool z0 = !mask_bit(MASK, 0), z1 = !mask_bit(MASK, 1);
    const cplx a = x[0], b = x[1];
    x[0] = zadd(z0, z1, a, b); x[1] = zsub(z0, z1, a, b);
  }
};

template <int S, unsigned MASK> struct DftM<4, S, MASK> {
  static __device__ __forceinline__ void run(cplx* x)
  {
    constexpr bool z0 = !mask_bit(MASK, 0), z1 = !mask_bit(MASK, 1), z2 = !mask_bit(MASK, 2), z3 = !mask_bit(MASK, 3);
    constexpr bool ze = z0 && z2, zo = z1 && z3;
    const cplx t0 = zadd(z0, z2, x[0], x[2]), t1 = zsub(z0, z2, x[0], x[2]);
    const cplx t2 = zadd(z1, z3, x[1], x[3]), t3 = mul_i<S>(zsub(z1, z3, x[1], x[3]));
    x[0] = zadd(ze, zo, t0, t2); x[2] = zsub(ze, zo, t0, t2);
    x[1] = zadd(ze, zo, t1, t3); x[3] = zsub(ze, zo, t1, t3);
  }
};

// odd prime radix with zero inputs: pair x[k], x[P-k]
template <int P, int S, unsigned MASK> struct DftPrimeM {
  static __device__ __forceinline__ void run(cplx* x)
  {
    constexpr int H = (P - 1) / 2;
    constexpr bool z0 = !mask_bit(MASK, 0);
    cplx a[H], b[H];
    bool za[H];
#pragma unroll
    for (int k = 1; k <= H; k++) {
      const bool zp = !mask_bit(MASK, k), zm = !mask_bit(MASK, P - k);
      a[k - 1] = zadd(zp, zm, x[k], x[P - k]);
      b[k - 1] = zsub(zp, zm, x[k], x[P - k]);
      za[k - 1] = zp && zm;
    }
    const cplx x0 = z0 ? make_double2(0.0, 0.0) : x[0];
    {
      cplx s0 = x0; bool zs = z0;
#pragma unroll
      for (int k = 0; k < H; k++) { s0 = zadd(zs, za[k], s0, a[k]); zs = zs && za[k]; }
      x[0] = s0;
    }
#pragma unroll
    for (int m = 1; m <= H; m++) {
      double re = x0.x, im = x0.y, dr = 0.0, di = 0.0;
      bool zr = z0, zd = true;
#pragma unroll
      for (int k = 1; k <= H; k++) {
        if (za[k - 1]) continue;
        const int e = (k * m) % P;
        const double c = Roots<P>::c(e), s = Roots<P>::s(e);
        if (zr) { re = c * a[k - 1].x; im = c * a[k - 1].y; zr = false; }
        else { re += c * a[k - 1].x; im += c * a[k - 1].y; }
        if (zd) { dr = s * b[k - 1].x; di = s * b[k - 1].y; zd = false; }
        else { dr += s * b[k - 1].x; di += s * b[k - 1].y; }
      }
      if (zd) { x[m] = make_double2(re, im); x[P - m] = make_double2(re, im); }
      else {
        x[m] = make_double2(re - S * di, im + S * dr);
        x[P - m] = make_double2(re + S * di, im - S * dr);
      }
    }
  }
};
template <int S, unsigned MASK> struct DftM<3, S, MASK> : DftPrimeM<3, S, MASK> {};
template <int S, unsigned MASK> struct DftM<5, S, MASK> : DftPrimeM<5, S, MASK> {};
template <int S, unsigned MASK> struct DftM<7, S, MASK> : DftPrimeM<7, S, MASK> {};
template <int S, unsigned MASK> struct DftM<11, S, MASK> : DftPrimeM<11, S, MASK> {};

// composite radix R = A*B (index maps as DftComposite): stage 1 over j1 for each j2 with the sub-mask of column j2,
// stage 2 over j2 with the mask of the columns that were not entirely zero
template <int A, int B, int S, unsigned MASK> struct DftCompositeM {
  static constexpr int R = A * B;
  static __host__ __device__ constexpr unsigned submask(int j2)
  {
    unsigned m = 0;
    for (int j1 = 0; j1 < A; j1++) if (mask_bit(MASK, B * j1 + j2)) m |= 1u << j1;
    return m;
  }
  static __host__ __device__ constexpr unsigned mask2()
  {
    unsigned m = 0;
    for (int j2 = 0; j2 < B; j2++) if (submask(j2) != 0) m |= 1u << j2;
    return m;
  }
  template <int J2> static __device__ __forceinline__ void stage1(const cplx* x, cplx* y)
  {
    constexpr unsigned sub = submask(J2);
    if constexpr (sub != 0) {
      cplx t[A];
#pragma unroll
      for (int j1 = 0; j1 < A; j1++) t[j1] = x[B * j1 + J2];
      DftM<A, S, sub>::run(t);
#pragma unroll
      for (int k1 = 0; k1 < A; k1++) y[k1 * B + J2] = mul_root<R, S>(t[k1], J2 * k1);
    }
    if constexpr (J2 + 1 < B) stage1<J2 + 1>(x, y);
  }
  static __device__ __forceinline__ void run(cplx* x)
  {
    cplx y[R];
    stage1<0>(x, y);
    constexpr unsigned m2 = mask2();
#pragma unroll
    for (int k1 = 0; k1 < A; k1++) {
      cplx t[B];
#pragma unroll
      for (int j2 = 0; j2 < B; j2++) if (mask_bit(m2, j2)) t[j2] = y[k1 * B + j2];
      DftM<B, S, m2>::run(t);
#pragma unroll
      for (int k2 = 0; k2 < B; k2++) x[k1 + A * k2] = t[k2];
    }
  }
};
template <int S, unsigned MASK> struct DftM<8, S, MASK> : DftCompositeM<2, 4, S, MASK> {};
template <int S, unsigned MASK> struct DftM<9, S, MASK> : DftCompositeM<3, 3, S, MASK> {};
template <int S, unsigned MASK> struct DftM<16, S, MASK> : DftCompositeM<4, 4, S, MASK> {};

}  // namespace qb200
