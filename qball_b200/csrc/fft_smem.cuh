// qball_b200/csrc/fft_smem.cuh
// Shared-memory multi-line complex-double FFT engine used by every transform kernel.
//
// A CTA holds `nlines` lines of length n in shared memory (element (line,j) at sm[lineoff(line) + j*estride]) and
// transforms all of them in place, in natural order, with the whole CTA:
//   * n = r[0]*r[1]*...*r[nf-1], radices from {16,8,4,2,9,3,5,7,11}
//   * steps 0..nf-2 are decimation-in-frequency passes: a task loads r elements at stride len/r into registers, does
//     the radix-r butterfly (fft_radix.cuh), applies the inter-step twiddles and stores to the SAME addresses (so no
//     barrier inside a step, one __syncthreads between steps);
//   * the last step reads runs of r[nf-1] adjacent elements and stores them to their digit-reversed (= natural)
//     positions; it walks the lines in rounds so one barrier separates a round's loads from its stores.
// The per-radix task loops are __noinline__ so that each gets its own register allocation (radix 16 needs ~112
// registers, radix 3 needs 32) instead of every kernel paying for the union of all inlined radices.
// Tasks are distributed line-fastest over the threads: with an odd line pitch (x rows) or unit line stride (y/z
// columns staged x-fastest) every quarter-warp touches 8 distinct 16-byte bank groups -> conflict-free LDS/STS.128.
#pragma once
#include "fft_radix.cuh"
#include "fft_desc.h"

namespace qb200 {

// where line `l` starts: lines may be split in two blocks (kept rows [0,nt) and [np1-nt,np1) of a plane)
struct LineMap {
  int lstride;   // elements between consecutive lines
  int lsplit;    // lines >= lsplit are shifted by lskip lines
  int lskip;
  __device__ __forceinline__ int off(int l) const { return (l < lsplit ? l : l + lskip) * lstride; }
};

// division by a loop-invariant positive divisor without the ~40-instruction integer divide: float reciprocal + fix-up
// (exact for 0 <= n < 2^24)
struct FastDiv {
  int d;
  float r;
  __device__ __forceinline__ explicit FastDiv(int dd) : d(dd), r(1.0f / (float)dd) {}
  __device__ __forceinline__ int div(int n, int& rem) const
  {
    int q = __float2int_rz(__int2float_rn(n) * r);
    rem = n - q * d;
    if (rem < 0) { q--; rem += d; }
    else if (rem >= d) { q++; rem -= d; }
    return q;
  }
};

template <int R, int S>
__device__ __noinline__ void dif_tasks(cplx* sm, int nlines, LineMap lm, int estride, int n, int len,
                                          const cplx* __restrict__ tw)
{
  const int m = len / R;
  const int per_line = n / R;                 // (n/len) segments * m offsets
  const int ntask = nlines * per_line;
  const int twmul = n / len;
  const FastDiv dl(nlines), dm(m);
  for (int task = threadIdx.x; task < ntask; task += blockDim.x) {
    int line, t;
    const int q = dl.div(task, line);
    const int seg = dm.div(q, t);
    cplx* p = sm + lm.off(line) + (seg * len + t) * estride;
    const int step = m * estride;
    cplx x[R];
#pragma unroll
    for (int k = 0; k < R; k++) x[k] = p[k * step];
    Dft<R, S>::run(x);
    p[0] = x[0];
    const int tws = t * twmul;
#pragma unroll
    for (int k = 1; k < R; k++) {
      const cplx w = tw[k * tws];
      p[k * step] = cmul_s<S>(x[k], w.x, w.y);
    }
  }
}

template <int R, int S>
__device__ __noinline__ void last_tasks(cplx* sm, int nlines, LineMap lm, int estride, const FftDesc& d)
{
  const int n = d.n;
  const int tpl = n / R;                       // tasks per line
  int lpr = blockDim.x / tpl;                  // lines per round
  if (lpr < 1) lpr = 1;                        // (tpl > blockDim.x is rejected on the host)
  if (lpr > nlines) lpr = nlines;
  int lr;
  const int u = FastDiv(lpr).div(threadIdx.x, lr);
  // natural position of run u: digits of u (most significant = first radix) reversed
  int rev = 0;
  {
    int rem = u, mul = 1, div = tpl;
    for (int i = 0; i < d.nf - 1; i++) {
      div /= d.r[i];
      const int dig = rem / div;
      rem -= dig * div;
      rev += dig * mul;
      mul *= d.r[i];
    }
  }
  // inactive threads (u >= tpl, or past the last line) work on clamped indices and simply do not store: keeping the
  // loads and the butterfly unconditional keeps x[] in registers across the barrier
  const int uc = min(u, tpl - 1);
  for (int line0 = 0; line0 < nlines; line0 += lpr) {
    const int line = line0 + lr;
    const bool act = (u < tpl) && (line < nlines);
    cplx x[R];
    cplx* base = sm + lm.off(min(line, nlines - 1));
#pragma unroll
    for (int k = 0; k < R; k++) x[k] = base[(uc * R + k) * estride];
    Dft<R, S>::run(x);
    __syncthreads();
    if (act) {
#pragma unroll
      for (int k = 0; k < R; k++) base[(rev + k * tpl) * estride] = x[k];
    }
  }
}

template <int S>
__device__ __forceinline__ void dif_step(int r, cplx* sm, int nlines, LineMap lm, int estride, int n, int len,
                                         const cplx* __restrict__ tw)
{
  switch (r) {
    case 16: dif_tasks<16, S>(sm, nlines, lm, estride, n, len, tw); break;
    case 8: dif_tasks<8, S>(sm, nlines, lm, estride, n, len, tw); break;
    case 4: dif_tasks<4, S>(sm, nlines, lm, estride, n, len, tw); break;
    case 2: dif_tasks<2, S>(sm, nlines, lm, estride, n, len, tw); break;
    case 9: dif_tasks<9, S>(sm, nlines, lm, estride, n, len, tw); break;
    case 3: dif_tasks<3, S>(sm, nlines, lm, estride, n, len, tw); break;
    case 5: dif_tasks<5, S>(sm, nlines, lm, estride, n, len, tw); break;
    case 7: dif_tasks<7, S>(sm, nlines, lm, estride, n, len, tw); break;
    case 11: dif_tasks<11, S>(sm, nlines, lm, estride, n, len, tw); break;
    default: break;
  }
}

template <int S>
__device__ __forceinline__ void last_step(int r, cplx* sm, int nlines, LineMap lm, int estride, const FftDesc& d)
{
  switch (r) {
    case 16: last_tasks<16, S>(sm, nlines, lm, estride, d); break;
    case 8: last_tasks<8, S>(sm, nlines, lm, estride, d); break;
    case 4: last_tasks<4, S>(sm, nlines, lm, estride, d); break;
    case 2: last_tasks<2, S>(sm, nlines, lm, estride, d); break;
    case 9: last_tasks<9, S>(sm, nlines, lm, estride, d); break;
    case 3: last_tasks<3, S>(sm, nlines, lm, estride, d); break;
    case 5: last_tasks<5, S>(sm, nlines, lm, estride, d); break;
    case 7: last_tasks<7, S>(sm, nlines, lm, estride, d); break;
    case 11: last_tasks<11, S>(sm, nlines, lm, estride, d); break;
    default: break;
  }
}

// In-place FFT of all lines.  Caller guarantees the data is visible (a __syncthreads before the call); on return all
// results are visible to the whole CTA (trailing __syncthreads).  tw[j] = (cos, sin)(2 pi j / n).
template <int S>
__device__ __forceinline__ void fft_lines(cplx* sm, int nlines, LineMap lm, int estride, const FftDesc& d,
                                          const cplx* __restrict__ tw)
{
  int len = d.n;
  for (int s = 0; s < d.nf - 1; s++) {
    dif_step<S>(d.r[s], sm, nlines, lm, estride, d.n, len, tw);
    __syncthreads();
    len /= d.r[s];
  }
  last_step<S>(d.r[d.nf - 1], sm, nlines, lm, estride, d);
  __syncthreads();
}

}  // namespace qb200
