// qball_b200/csrc/fft_group.cuh
// Group-synchronous shared-memory FFT passes (second-generation engine, used by the plane-fused xy kernels and the
// z-column kernels).
//
// Differences from fft_smem.cuh (which the split path still uses):
//   * a pass works on a BLOCK of up to 8 independent lines and is executed by a GROUP of threads (a few warps) that
//     synchronises with a named barrier (bar.sync id, nthreads) -- groups of one CTA run different blocks out of
//     phase, so one group's shared-memory traffic overlaps another group's FP64 butterflies; there is no CTA-wide
//     barrier inside a transform;
//   * no un-permuting pass: transforms come in transposed pairs.  A decimation-in-frequency (DIF) transform maps
//     natural order -> digit-reversed order in place; the transposed network (decimation in time, DIT: twiddle, then
//     butterfly, passes in reverse order) maps digit-reversed -> natural in place.  Pointwise work between a DIF and
//     a DIT transform happens in digit-reversed order (the "mid" pass: last DIF butterfly, v(r) multiply or |psi|^2
//     accumulate, first DIT butterfly -- all in registers), and scatter/gather tables absorb the permutation at the
//     sphere side.  Position q = d0*(n/r0) + d1*(n/(r0 r1)) + ... + d_{f-1} holds natural index
//     d0 + r0*(d1 + r1*(d2 + ...)).
//   * twiddles come from a packed per-pass table (fft_desc.h): the r-1 factors of a task are contiguous;
//   * the first DIF pass / last DIT pass can skip rows known to be zero / not needed (the reference's ntrans0 pruning,
//     FourierTransform.cc:202, 772-819, applied to the y direction as well).
#pragma once
#include "fft_smem.cuh"

namespace qb200 {

struct Grp {
  int tid, nthr, bar;
  __device__ __forceinline__ void sync() const { asm volatile("bar.sync %0, %1;" ::"r"(bar), "r"(nthr) : "memory"); }
};

// kept index range of a pruned direction: [0,ksplit) and [ksplit+kskip, n)
struct Keep {
  int ksplit, kskip;
  __device__ __forceinline__ bool kept(int j) const { return j < ksplit || j >= ksplit + kskip; }
};

#define QB200_BLOCK_LINES 8   // lines per y block (8 x 16 B = one 128-byte shared-memory wavefront per quarter warp)

// One radix-R pass over `nlines` independent lines.  Element (line, j) lives at base[lm.off(line) + j*estride]; tasks
// are distributed line-fastest over the group's threads, `nlpad` >= nlines being the line count the task index is
// decomposed with (8 for a column block: shifts; lanes of missing lines idle).
//   DIT == false: butterfly, then multiply output k by w_len^(k t)      (decimation in frequency)
//   DIT == true : multiply input k by w_len^(k t), then butterfly       (transposed network)
//   prune (only meaningful for the pass with len == n): DIF skips loads of rows outside `kp` (zeros), DIT skips
//   stores to rows outside `kp`.
template <int R, int S, bool DIT>
__device__ __noinline__ void radix_pass(Grp g, cplx* base, int nlines, int nlpad, LineMap lm, int estride, int n, int len,
                                        const cplx* __restrict__ tw, bool prune, Keep kp)
{
  const int m = len / R;
  const int ntask = (n / R) * nlpad;
  const FastDiv dm(m), dl(nlpad);
  const int step = m * estride;
  for (int task = g.tid; task < ntask; task += g.nthr) {
    int line, t;
    const int q = dl.div(task, line);
    if (line >= nlines) continue;
    const int seg = dm.div(q, t);
    cplx* p = base + lm.off(line) + (seg * len + t) * estride;
    cplx x[R];
    if (!DIT && prune) {
#pragma unroll
      for (int k = 0; k < R; k++) x[k] = kp.kept(t + k * m) ? p[k * step] : make_double2(0.0, 0.0);
    } else {
#pragma unroll
      for (int k = 0; k < R; k++) x[k] = p[k * step];
    }
    const cplx* twt = tw + t * (R - 1) - 1;      // packed: entry (t, k) at t*(R-1) + k-1
    if (DIT && m > 1) {
#pragma unroll
      for (int k = 1; k < R; k++) {
        const cplx w = twt[k];
        x[k] = cmul_s<S>(x[k], w.x, w.y);
      }
    }
    Dft<R, S>::run(x);
    if (!DIT && m > 1) {
#pragma unroll
      for (int k = 1; k < R; k++) {
        const cplx w = twt[k];
        x[k] = cmul_s<S>(x[k], w.x, w.y);
      }
    }
    if (DIT && prune) {
#pragma unroll
      for (int k = 0; k < R; k++) if (kp.kept(t + k * m)) p[k * step] = x[k];
    } else {
#pragma unroll
      for (int k = 0; k < R; k++) p[k * step] = x[k];
    }
  }
}

template <int S, bool DIT>
__device__ __forceinline__ void radix_pass_any(int r, Grp g, cplx* base, int nlines, int nlpad, LineMap lm, int estride, int n,
                                               int len, const cplx* __restrict__ tw, bool prune, Keep kp)
{
  switch (r) {
    case 16: radix_pass<16, S, DIT>(g, base, nlines, nlpad, lm, estride, n, len, tw, prune, kp); break;
    case 8: radix_pass<8, S, DIT>(g, base, nlines, nlpad, lm, estride, n, len, tw, prune, kp); break;
    case 4: radix_pass<4, S, DIT>(g, base, nlines, nlpad, lm, estride, n, len, tw, prune, kp); break;
    case 2: radix_pass<2, S, DIT>(g, base, nlines, nlpad, lm, estride, n, len, tw, prune, kp); break;
    case 9: radix_pass<9, S, DIT>(g, base, nlines, nlpad, lm, estride, n, len, tw, prune, kp); break;
    case 3: radix_pass<3, S, DIT>(g, base, nlines, nlpad, lm, estride, n, len, tw, prune, kp); break;
    case 5: radix_pass<5, S, DIT>(g, base, nlines, nlpad, lm, estride, n, len, tw, prune, kp); break;
    case 7: radix_pass<7, S, DIT>(g, base, nlines, nlpad, lm, estride, n, len, tw, prune, kp); break;
    case 11: radix_pass<11, S, DIT>(g, base, nlines, nlpad, lm, estride, n, len, tw, prune, kp); break;
    default: break;
  }
}

// natural -> digit-reversed, in place, passes first_pass .. first_pass+npass-1; a group barrier separates consecutive
// passes (none at the end)
template <int S>
__device__ __forceinline__ void fft_block_dif(Grp g, cplx* base, int nlines, int nlpad, LineMap lm, int estride, const FftDesc& d,
                                              const cplx* __restrict__ tw, int first_pass, int npass, bool prune, Keep kp)
{
  for (int s = first_pass; s < first_pass + npass; s++) {
    if (s > first_pass) g.sync();
    radix_pass_any<S, false>(d.r[s], g, base, nlines, nlpad, lm, estride, d.n, d.len[s], tw + d.twoff[s], prune && s == 0, kp);
  }
}

// digit-reversed -> natural, in place: passes last_pass, last_pass-1, ..., 0
template <int S>
__device__ __forceinline__ void fft_block_dit(Grp g, cplx* base, int nlines, int nlpad, LineMap lm, int estride, const FftDesc& d,
                                              const cplx* __restrict__ tw, int last_pass, bool prune, Keep kp)
{
  for (int s = last_pass; s >= 0; s--) {
    if (s < last_pass) g.sync();
    radix_pass_any<S, true>(d.r[s], g, base, nlines, nlpad, lm, estride, d.n, d.len[s], tw + d.twoff[s], prune && s == 0, kp);
  }
}

}  // namespace qb200
