// qball_b200/csrc/fft_desc.h
// Factorisation of a grid length into FFT radices, usable at compile time (shape-specialised kernels) and at run time
// (generic kernels, host tables) -- one rule, so that the digit-reversal tables built on the host match what the
// kernels do.  Lengths are those Basis::factorizable allows (/root/reference/src/qball/Basis.cc:126-147):
// n = 2^a 3^b 5^c 7^d 11^e.
#pragma once

#ifdef __CUDACC__
#define QB200_HD __host__ __device__
#else
#define QB200_HD
#endif

namespace qb200 {

#define QB200_MAXF 8

struct FftDesc {
  int n;
  int nf;
  int r[QB200_MAXF];
  int len[QB200_MAXF];     // sub-transform length entering pass s: n / (r[0]*...*r[s-1])
  int twoff[QB200_MAXF];   // offset of pass s in the packed twiddle table: entry (t, k) at twoff + t*(r-1) + k-1, t < len/r, 1 <= k < r
  int twsize;              // packed twiddle table size (>= 1)
};

// radices descending (the twiddled passes take the big radices), the smallest last.  nf == 0: not factorisable.
QB200_HD constexpr FftDesc make_fft_desc(int n)
{
  FftDesc d = {};
  d.n = n; d.nf = 0; d.twsize = 1;
  if (n < 1) return d;
  int f[32] = {};
  int nf = 0;
  int m = n;
  const int primes[3] = { 11, 7, 5 };
  for (int i = 0; i < 3; i++) while (m % primes[i] == 0 && nf < 32) { f[nf++] = primes[i]; m /= primes[i]; }
  while (m % 9 == 0 && nf < 32) { f[nf++] = 9; m /= 9; }
  while (m % 3 == 0 && nf < 32) { f[nf++] = 3; m /= 3; }
  while (m % 16 == 0 && nf < 32) { f[nf++] = 16; m /= 16; }
  if (m % 8 == 0) { f[nf++] = 8; m /= 8; }
  if (m % 4 == 0) { f[nf++] = 4; m /= 4; }
  if (m % 2 == 0) { f[nf++] = 2; m /= 2; }
  if (m != 1) return d;
  if (nf == 0) f[nf++] = 1;
  if (nf > QB200_MAXF) return d;
  for (int i = 1; i < nf; i++) {                     // insertion sort, descending
    const int v = f[i];
    int j = i - 1;
    while (j >= 0 && f[j] < v) { f[j + 1] = f[j]; j--; }
    f[j + 1] = v;
  }
  // the last pass of the first-generation engine needs n/r tasks per line in one round of <= 256 threads
  for (int i = nf - 1; i >= 0; i--)
    if (n / f[i] <= 256) {
      const int r = f[i];
      for (int j = i; j + 1 < nf; j++) f[j] = f[j + 1];
      f[nf - 1] = r;
      break;
    }
  d.nf = nf;
  int len = n, off = 0;
  for (int i = 0; i < nf; i++) {
    d.r[i] = f[i]; d.len[i] = len;
    d.twoff[i] = off;
    const int mm = len / f[i];
    if (mm > 1) off += mm * (f[i] - 1);
    len /= f[i];
  }
  d.twsize = off > 0 ? off : 1;
  return d;
}

// natural index held by in-place position q after a DIF transform with d's radices
QB200_HD constexpr int digit_reverse(const FftDesc& d, int q)
{
  int div = d.n, nat = 0, mul = 1;
  for (int i = 0; i < d.nf; i++) {
    div /= d.r[i];
    const int dig = q / div;
    q -= dig * div;
    nat += dig * mul;
    mul *= d.r[i];
  }
  return nat;
}

}  // namespace qb200
