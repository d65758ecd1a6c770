// qball_b200/csrc/fft_radix.cuh
// Register-resident complex-double DFT butterflies for the radices that factor every grid length the reference
// allows (Basis::factorizable, /root/reference/src/qball/Basis.cc:126-147: n = 2^a 3^{<=2} 5^{<=1} 7^{<=1} 11^{<=1}).
// Sign convention as the reference's FFTW plans: S=+1 for backward (e^{+iGr}), S=-1 for forward
// (FourierTransform.cc:1516-1612).  Constant tables are generated to 20 digits (mpmath) -- FP64 twiddles, no
// fast-math sincos -- so fwd(bwd(c)) holds 1e-10 with a wide margin.
#pragma once
#include <cuda_runtime.h>

namespace qb200 {

typedef double2 cplx;

__device__ __forceinline__ cplx cadd(cplx a, cplx b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ cplx csub(cplx a, cplx b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ cplx cmul(cplx a, cplx b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
// a * (wr + i*S*wi)
template <int S> __device__ __forceinline__ cplx cmul_s(cplx a, double wr, double wi)
{
  return make_double2(a.x * wr - S * (a.y * wi), a.y * wr + S * (a.x * wi));
}
// multiply by S*i
template <int S> __device__ __forceinline__ cplx mul_i(cplx a) { return make_double2(-S * a.y, S * a.x); }

template <int N> struct Roots;
template<> struct Roots<3> {
  static __device__ __forceinline__ double c(int k) { constexpr double t[3] = { 1.0, -0.5, -0.5 }; return t[k]; }
  static __device__ __forceinline__ double s(int k) { constexpr double t[3] = { 0.0, 0.86602540378443864676, -0.86602540378443864676 }; return t[k]; }
};
template<> struct Roots<5> {
  static __device__ __forceinline__ double c(int k) { constexpr double t[5] = { 1.0, 0.3090169943749474241, -0.8090169943749474241, -0.8090169943749474241, 0.3090169943749474241 }; return t[k]; }
  static __device__ __forceinline__ double s(int k) { constexpr double t[5] = { 0.0, 0.95105651629515357212, 0.58778525229247312917, -0.58778525229247312917, -0.95105651629515357212 }; return t[k]; }
};
template<> struct Roots<7> {
  static __device__ __forceinline__ double c(int k) { constexpr double t[7] = { 1.0, 0.62348980185873353053, -0.22252093395631440429, -0.90096886790241912624, -0.90096886790241912624, -0.22252093395631440429, 0.62348980185873353053 }; return t[k]; }
  static __device__ __forceinline__ double s(int k) { constexpr double t[7] = { 0.0, 0.78183148246802980871, 0.97492791218182360702, 0.43388373911755812048, -0.43388373911755812048, -0.97492791218182360702, -0.78183148246802980871 }; return t[k]; }
};
template<> struct Roots<8> {
  static __device__ __forceinline__ double c(int k) { constexpr double t[8] = { 1.0, 0.7071067811865475244, 0.0, -0.7071067811865475244, -1.0, -0.7071067811865475244, 0.0, 0.7071067811865475244 }; return t[k]; }
  static __device__ __forceinline__ double s(int k) { constexpr double t[8] = { 0.0, 0.7071067811865475244, 1.0, 0.7071067811865475244, 0.0, -0.7071067811865475244, -1.0, -0.7071067811865475244 }; return t[k]; }
};
template<> struct Roots<9> {
  static __device__ __forceinline__ double c(int k) { constexpr double t[9] = { 1.0, 0.7660444431189780352, 0.17364817766693034885, -0.5, -0.93969262078590838405, -0.93969262078590838405, -0.5, 0.17364817766693034885, 0.7660444431189780352 }; return t[k]; }
  static __device__ __forceinline__ double s(int k) { constexpr double t[9] = { 0.0, 0.64278760968653932632, 0.98480775301220805937, 0.86602540378443864676, 0.34202014332566873304, -0.34202014332566873304, -0.86602540378443864676, -0.98480775301220805937, -0.64278760968653932632 }; return t[k]; }
};
template<> struct Roots<11> {
  static __device__ __forceinline__ double c(int k) { constexpr double t[11] = { 1.0, 0.84125353283118116886, 0.41541501300188642553, -0.14231483827328514044, -0.65486073394528506406, -0.95949297361449738989, -0.95949297361449738989, -0.65486073394528506406, -0.14231483827328514044, 0.41541501300188642553, 0.84125353283118116886 }; return t[k]; }
  static __device__ __forceinline__ double s(int k) { constexpr double t[11] = { 0.0, 0.54064081745559758211, 0.90963199535451837141, 0.98982144188093273238, 0.75574957435425828377, 0.28173255684142969771, -0.28173255684142969771, -0.75574957435425828377, -0.98982144188093273238, -0.90963199535451837141, -0.54064081745559758211 }; return t[k]; }
};
template<> struct Roots<14> {
  static __device__ __forceinline__ double c(int k) { constexpr double t[14] = { 1.0, 0.90096886790241912624, 0.62348980185873353053, 0.22252093395631440429, -0.22252093395631440429, -0.62348980185873353053, -0.90096886790241912624, -1.0, -0.90096886790241912624, -0.62348980185873353053, -0.22252093395631440429, 0.22252093395631440429, 0.62348980185873353053, 0.90096886790241912624 }; return t[k]; }
  static __device__ __forceinline__ double s(int k) { constexpr double t[14] = { 0.0, 0.43388373911755812048, 0.78183148246802980871, 0.97492791218182360702, 0.97492791218182360702, 0.78183148246802980871, 0.43388373911755812048, 0.0, -0.43388373911755812048, -0.78183148246802980871, -0.97492791218182360702, -0.97492791218182360702, -0.78183148246802980871, -0.43388373911755812048 }; return t[k]; }
};
template<> struct Roots<16> {
  static __device__ __forceinline__ double c(int k) { constexpr double t[16] = { 1.0, 0.92387953251128675613, 0.7071067811865475244, 0.38268343236508977173, 0.0, -0.38268343236508977173, -0.7071067811865475244, -0.92387953251128675613, -1.0, -0.92387953251128675613, -0.7071067811865475244, -0.38268343236508977173, 0.0, 0.38268343236508977173, 0.7071067811865475244, 0.92387953251128675613 }; return t[k]; }
  static __device__ __forceinline__ double s(int k) { constexpr double t[16] = { 0.0, 0.38268343236508977173, 0.7071067811865475244, 0.92387953251128675613, 1.0, 0.92387953251128675613, 0.7071067811865475244, 0.38268343236508977173, 0.0, -0.38268343236508977173, -0.7071067811865475244, -0.92387953251128675613, -1.0, -0.92387953251128675613, -0.7071067811865475244, -0.38268343236508977173 }; return t[k]; }
};

// multiply x by w_R^m (sign S) with m known after unrolling: trivial cases cost nothing
template <int R, int S> __device__ __forceinline__ cplx mul_root(cplx x, int m)
{
  m %= R;
  if (m == 0) return x;
  if (2 * m == R) return make_double2(-x.x, -x.y);
  if (4 * m == R) return mul_i<S>(x);
  if (4 * m == 3 * R) return mul_i<-S>(x);
  return cmul_s<S>(x, Roots<R>::c(m), Roots<R>::s(m));
}

template <int R, int S> struct Dft;

template <int S> struct Dft<1, S> { static __device__ __forceinline__ void run(cplx*) {} };

template <int S> struct Dft<2, S> {
  static __device__ __forceinline__ void run(cplx* x)
  {
    cplx a = x[0], b = x[1];
    x[0] = cadd(a, b); x[1] = csub(a, b);
  }
};

template <int S> struct Dft<4, S> {
  static __device__ __forceinline__ void run(cplx* x)
  {
    cplx t0 = cadd(x[0], x[2]), t1 = csub(x[0], x[2]), t2 = cadd(x[1], x[3]), t3 = mul_i<S>(csub(x[1], x[3]));
    x[0] = cadd(t0, t2); x[2] = csub(t0, t2); x[1] = cadd(t1, t3); x[3] = csub(t1, t3);
  }
};

// odd prime radix: pair x[k], x[P-k]
template <int P, int S> struct DftPrime {
  static __device__ __forceinline__ void run(cplx* x)
  {
    constexpr int H = (P - 1) / 2;
    cplx a[H], b[H];
#pragma unroll
    for (int k = 1; k <= H; k++) { a[k - 1] = cadd(x[k], x[P - k]); b[k - 1] = csub(x[k], x[P - k]); }
    cplx x0 = x[0];
    cplx s0 = x0;
#pragma unroll
    for (int k = 0; k < H; k++) s0 = cadd(s0, a[k]);
    x[0] = s0;
#pragma unroll
    for (int m = 1; m <= H; m++) {
      double re = x0.x, im = x0.y, dr = 0.0, di = 0.0;
#pragma unroll
      for (int k = 1; k <= H; k++) {
        const int e = (k * m) % P;
        const double c = Roots<P>::c(e), s = Roots<P>::s(e);
        re += c * a[k - 1].x; im += c * a[k - 1].y;
        dr += s * b[k - 1].x; di += s * b[k - 1].y;
      }
      // X[m] = (re,im) + S*i*(dr,di) ; X[P-m] = (re,im) - S*i*(dr,di)
      x[m] = make_double2(re - S * di, im + S * dr);
      x[P - m] = make_double2(re + S * di, im - S * dr);
    }
  }
};
template <int S> struct Dft<3, S> : DftPrime<3, S> {};
template <int S> struct Dft<5, S> : DftPrime<5, S> {};
template <int S> struct Dft<7, S> : DftPrime<7, S> {};
template <int S> struct Dft<11, S> : DftPrime<11, S> {};

// composite radix R = A*B:  j = B*j1 + j2, k = k1 + A*k2
//   X[k1 + A*k2] = sum_{j2} w_B^{j2 k2} [ w_R^{j2 k1} sum_{j1} w_A^{j1 k1} x[B*j1 + j2] ]
template <int A, int B, int S> struct DftComposite {
  static __device__ __forceinline__ void run(cplx* x)
  {
    constexpr int R = A * B;
    cplx y[R];
#pragma unroll
    for (int j2 = 0; j2 < B; j2++) {
      cplx t[A];
#pragma unroll
      for (int j1 = 0; j1 < A; j1++) t[j1] = x[B * j1 + j2];
      Dft<A, S>::run(t);
#pragma unroll
      for (int k1 = 0; k1 < A; k1++) y[k1 * B + j2] = mul_root<R, S>(t[k1], j2 * k1);
    }
#pragma unroll
    for (int k1 = 0; k1 < A; k1++) {
      cplx t[B];
#pragma unroll
      for (int j2 = 0; j2 < B; j2++) t[j2] = y[k1 * B + j2];
      Dft<B, S>::run(t);
#pragma unroll
      for (int k2 = 0; k2 < B; k2++) x[k1 + A * k2] = t[k2];
    }
  }
};
template <int S> struct Dft<8, S> : DftComposite<2, 4, S> {};
template <int S> struct Dft<9, S> : DftComposite<3, 3, S> {};
template <int S> struct Dft<16, S> : DftComposite<4, 4, S> {};
template <int S> struct Dft<14, S> : DftComposite<2, 7, S> {};   // thread-private passes of the tensor-memory kernels (126 = 9 x 14)

}  // namespace qb200
