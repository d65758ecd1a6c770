// qball_b200/csrc/plane_tmem.cuh
// k_plane_t<OP, SH>: the plane-fused xy stage with the y direction in TENSOR MEMORY (Blackwell TMEM, tcgen05.ld/st) and the
// two directions running concurrently on different warps.  Same arithmetic as k_plane_s (FourierTransform.cc:761-974 backward,
// :1156-1361 forward, SlaterDet.cc:919-921 / :993-1031 for the pointwise work), different data flow:
//
//   * only the 2*ntrans0 kept rows of a plane live in shared memory (52 x 113 x 16 B = 94 KB instead of 202 KB), twice:
//     buffer A[i & 1] belongs to the i-th unit of the CTA;
//   * X warps (NXW of them) own the x direction: TMA bulk copy of the next unit's column values into a staging row
//     (cp.async.bulk + mbarrier), scatter, pruned x-DIT (compiled passes of plane_static.cuh), hand the buffer to the Y warps,
//     and for the unit before: pruned x-DIF, gather, store;
//   * Y warps (8 = two per TMEM lane quarter) own the y direction, ONE THREAD PER COLUMN: the 112-point column sits in the
//     thread's own TMEM lane (112 x 16 B = 448 of the 512 32-bit columns), so the radix-16 and radix-7 passes of the round
//     trip  y-transform(+1) -> v(r) multiply or |psi|^2 -> y-transform(-1)  need NO exchange between threads: inputs come
//     from the kept rows (lanes read consecutive x: conflict-free), every intermediate moves registers <-> TMEM, twiddles
//     are warp-uniform constants.  The two warps of a lane quarter split the butterflies of a pass (b = 0..3 | 4..6 for the
//     radix-16 passes, k1 split for the radix-7 pass) and meet at a named barrier of their own between passes (H psi: two warps
//     per quarter and 8 X warps; density, which has no way back: three per quarter and 4 X warps).
//
// The y direction was 2/3 of the shared-memory wavefronts of k_plane_s (three passes over the full 112 x 112 plane, plus
// per-thread twiddle loads); here it costs two reads/writes of the kept rows.  The shared-memory pipe (X warps) and the FP64
// pipe (mostly Y warps) work on different units at the same time instead of alternating between barriers.
//
// Index maps of the y direction (N = 112 = 16 x 7, S = +1 backward, -1 forward; W_n = exp(2 pi i / n)):
//   y = 7a + b, k = k1 + 16 k2:   X[k1 + 16 k2] = sum_b W_7^{S b k2} [ W_112^{S b k1} sum_a W_16^{S a k1} x[7a + b] ]
//   and transposed for the way back: Y[7a + b] = sum_k1 W_16^{-a k1} [ W_112^{-b k1} sum_k2 W_7^{-b k2} X[k1 + 16 k2] ].
//   TMEM slot (b, k1) of a lane = 32-bit columns 4*(16 b + k1) .. +3.
#pragma once
#include "plane_static.cuh"
#include "tmem_ops.cuh"
#include "async_ops.cuh"

namespace qb200 {

// c_ytw[16 b + k1] = W_112^{b k1} (cos, sin), filled by plane_t_setup (plane.cu)
__constant__ double2 c_ytw[7 * 16];
// c_yrow[8 k1 + k2] = (49 k1 + 64 k2) mod 112: the row of output (k1, k2) of the Good-Thomas passes (below), times np0
__constant__ int c_yrow[16 * 8];

// dynamic shared memory of k_plane_t (bytes); the same rule on the host (plane.cu)
template <class SH> QB200_HD constexpr size_t plane_t_smem(int nvec, int nzero)
{
  constexpr FftDesc FX = make_fft_desc(SH::NP0);
  size_t b = (size_t)((FX.twsize + 7) & ~7) * 16;                 // x twiddles
  b += 2 * (size_t)SH::NKEEP * SH::PITCH * 16;                   // A[0], A[1]
  b += (size_t)((nvec + 7) & ~7) * 16;                           // staging row
  b += (size_t)((nvec + 7) & ~7) * 2 + (size_t)((nzero + 7) & ~7) * 2;   // tpos, tzero (16-bit positions)
  return b;
}

enum { BAR_X = 1, BAR_FULL = 2, BAR_DONE = 4, BAR_PAIR = 6 };

// Good-Thomas (prime-factor) index maps of the thread-per-column y passes of k_plane_t: 16 and 7 are coprime, so with
//   y = (7a + 16b) mod 112,  k = (49 k1 + 64 k2) mod 112      (49 = 7 * (7^-1 mod 16), 64 = 16 * (16^-1 mod 7))
// W_112^{y k} = W_16^{a k1} W_7^{b k2} exactly: a 16 x 7 two-dimensional transform with NO twiddle factors between the passes
// (the Cooley-Tukey maps y = 7a + b, k = k1 + 16 k2 cost 105 + 96 complex multiplications and their table fetches per column
// and round trip).  Residue class b is a compile-time constant in passes 1 and 3, so the kept rows of a butterfly (its zero
// inputs / unused outputs) and their shared-memory offsets are constants too.
template <class SH, int B> struct YGT {
  static QB200_HD constexpr int y(int a) { return (7 * a + 16 * B) % 112; }
  static QB200_HD constexpr bool kept(int a) { return y(a) < SH::YSPLIT || y(a) >= SH::YSPLIT + SH::YSKIP; }
  static QB200_HD constexpr int row(int a) { return y(a) < SH::YSPLIT ? y(a) : y(a) - SH::YSKIP; }
  static QB200_HD constexpr unsigned mask() { unsigned m = 0; for (int a = 0; a < 16; a++) if (kept(a)) m |= 1u << a; return m; }
};
// pass 1 of class B: kept rows y(a) of the column -> 16-point transform over a -> TMEM slots (B, k1)
template <class SH, int B> __device__ __forceinline__ void y_pass1_gt(const cplx* __restrict__ Ax, uint32_t t0)
{
  typedef YGT<SH, B> M;
  cplx x[16];
#pragma unroll
  for (int a = 0; a < 16; a++) if (M::kept(a)) x[a] = Ax[M::row(a) * SH::PITCH];
  DftM<16, +1, M::mask()>::run(x);
  Tmem<16>::st(t0 + 64 * B, x);
}
// pass 3 of class B: slots (B, k1) -> 16-point transform over k1 -> the kept rows y(a)
template <class SH, int B> __device__ __forceinline__ void y_pass3_gt(cplx* __restrict__ Ax, uint32_t t0, bool act)
{
  typedef YGT<SH, B> M;
  cplx x[16];
  Tmem<16>::ld(x, t0 + 64 * B);
  Dft<16, -1>::run(x);
  if (act) {
#pragma unroll
    for (int a = 0; a < 16; a++) if (M::kept(a)) Ax[M::row(a) * SH::PITCH] = x[a];
  }
}

template <int OP, class SH, int NYW, int NXW>
__global__ void __launch_bounds__((NYW + NXW) * 32, 1) k_plane_t(const __grid_constant__ DevPlan P, cplx* __restrict__ zt, const double* __restrict__ v,
                                                                 double* __restrict__ rho_part, const double* __restrict__ fac, int nunits,
                                                                 int zero_imag)
{
  static_assert(OP == OP_HPSI || OP == OP_DENSITY, "k_plane_t: H psi and density only");
  static_assert(SH::NP1 == 112 && NYW % 4 == 0 && NYW >= 4 && NYW <= 16, "thread-per-column y passes are written for 112 = 16 x 7, MW = NYW/4 warps per TMEM lane quarter");
  static_assert(SH::NP0 <= 4 * 28, "28 columns per lane quarter");
  constexpr FftDesc FX = make_fft_desc(SH::NP0);
  constexpr int np0 = SH::NP0, np1 = SH::NP1, pitch = SH::PITCH, np01 = np0 * np1, NK = SH::NKEEP;
  constexpr int NYT = NYW * 32, NXT = NXW * 32, NT = NYT + NXT;
  constexpr int ABUF = NK * pitch;
  static_assert(2 * ABUF <= 65535, "16-bit plane positions");
  extern __shared__ __align__(128) unsigned char smraw[];
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(8) uint64_t mbar;
  const int nvec = P.nvec, nvp = (nvec + 7) & ~7, nzero = P.ntzero;
  cplx* tw0 = reinterpret_cast<cplx*>(smraw);
  cplx* A0 = tw0 + ((FX.twsize + 7) & ~7);
  cplx* stg = A0 + 2 * ABUF;
  unsigned short* tpos = reinterpret_cast<unsigned short*>(stg + nvp);
  unsigned short* tzero = tpos + nvp;
  // (the shuffle tells the compiler that the warp index -- and every role, loop bound and TMEM address derived from it -- is
  // warp-uniform: they live in uniform registers, no R2UR per tcgen05 instruction)
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
  const int z = blockIdx.x, G = gridDim.y;
  const size_t N = (size_t)np01 * P.np2;

  if (warp == 0) tmem_alloc512(&tmem_slot);
  if (tid == 32) mbar_init(&mbar, 1);
  for (int i = tid; i < FX.twsize; i += NT) tw0[i] = P.tw0p[i];
  for (int i = tid; i < nvec; i += NT) tpos[i] = P.tpos[i];
  for (int i = tid; i < nzero; i += NT) tzero[i] = P.tzero[i];
  for (int i = tid; i < 2 * ABUF; i += NT) A0[i] = make_double2(0.0, 0.0);
  tmem_fence_before();
  __syncthreads();
  tmem_fence_after();
  const uint32_t tbase = tmem_slot;

  auto next_unit = [&](int u) {
    u += G;
    if (OP == OP_DENSITY) while (u < nunits && !fac_active(P, fac, u)) u += G;
    return u;
  };
  const int first = next_unit((int)blockIdx.y - G);

  if (warp < NYW) {
    // ------------------------------------------------------------------------------------------ Y warps: one thread per column
    const int q = warp & 3, m = warp >> 2;
    const uint32_t t0 = tbase + ((uint32_t)(q * 32) << 16);
    const bool act = lane < 28 && 28 * q + lane < np0;
    const int xc = min(28 * q + lane, np0 - 1);
    constexpr int MW = NYW / 4;                    // warps that share a lane quarter split the butterflies of every pass
    const int blo = (7 * m) / MW, bhi = (7 * (m + 1)) / MW;
    const int klo = (16 * m) / MW, khi = (16 * (m + 1)) / MW;
    constexpr unsigned MASK = zmask(16, 7, SH::YSPLIT, SH::YSKIP);
    const double* vz = v + (size_t)z * np01 + xc;
    double* rz = rho_part + (size_t)blockIdx.y * N + (size_t)z * np01 + xc;
    int i = 0;
    for (int unit = first; unit < nunits; unit = next_unit(unit), i++) {
      cplx* A = A0 + (i & 1) * ABUF + xc;
      double facu = 0.0, facv = 0.0;
      if (OP == OP_DENSITY) { facu = fac_first(fac, unit); facv = fac_second(P, fac, unit); }
      bar_sync_n(BAR_FULL + (i & 1), NT);          // the X warps finished the x transform of this unit
      // pass 1: for each residue class b the 16-point transform over a of the kept rows y = (7a + 16b) mod 112 -> TMEM slots (b, .)
#define QB200_Y1(B) if (B >= blo && B < bhi) y_pass1_gt<SH, B>(A, t0);
      QB200_Y1(0) QB200_Y1(1) QB200_Y1(2) QB200_Y1(3) QB200_Y1(4) QB200_Y1(5) QB200_Y1(6)
#undef QB200_Y1
      tmem_wait_st();
      tmem_fence_before();
      bar_sync_n(BAR_PAIR + q, 32 * MW);
      tmem_fence_after();
      if (OP == OP_DENSITY) bar_arrive_n(BAR_DONE + (i & 1), NT);   // the kept rows are consumed: the buffer is free again
      // pass 2: for each k1 the 7-point transform over b -> psi(x, y, z) at y = (49 k1 + 64 k2) mod 112; pointwise work; way back
      // to slots (., k1).  (H psi: v(r) of the NEXT k1 is loaded while the current one is transformed; the loop is unrolled by two
      // with the two register sets swapping roles, so the prefetch costs no register moves)
      auto yrow = [](int k1, int k2) { return c_yrow[8 * k1 + k2]; };   // warp-uniform: ((49 k1 + 64 k2) mod 112) * np0
      auto pass2 = [&](int k1, const double (&vv)[7]) {
        cplx t[7];
        Tmem<1, 7>::ld(t, t0 + 4 * k1, 64);
        Dft<7, +1>::run(t);
        if (OP == OP_HPSI) {
          // (the odd tail of a real basis, whose imaginary part is dropped here -- SlaterDet.cc:1014-1023 -- runs k_plane_s)
#pragma unroll
          for (int k2 = 0; k2 < 7; k2++) { t[k2].x *= vv[k2]; t[k2].y *= vv[k2]; }
          Dft<7, -1>::run(t);
          Tmem<1, 7>::st2(t0 + 4 * k1, t, 64);
        } else {
          // fire-and-forget reductions at the L2, one owner per address (CTA (z, gy) owns plane z of partial gy), applied in
          // unit order: deterministic (as k_plane_s)
          if (act) {
#pragma unroll
            for (int k2 = 0; k2 < 7; k2++) {
              const double val = facu * t[k2].x * t[k2].x + facv * t[k2].y * t[k2].y;
              asm volatile("red.global.add.f64 [%0], %1;" ::"l"(rz + yrow(k1, k2)), "d"(val) : "memory");
            }
          }
        }
      };
      auto loadv = [&](int k1, double (&vv)[7]) {
        if (OP == OP_HPSI) {
#pragma unroll
          for (int k2 = 0; k2 < 7; k2++) vv[k2] = __ldg(vz + yrow(k1, k2));
        }
      };
      double va[7], vb[7];
      loadv(klo, va);
#pragma unroll 1
      for (int k1 = klo; k1 < khi; k1 += 2) {
        loadv(min(k1 + 1, khi - 1), vb);           // (the last iteration re-reads its own rows: no branch around the loads)
        pass2(k1, va);
        if (k1 + 1 < khi) {
          loadv(min(k1 + 2, khi - 1), va);
          pass2(k1 + 1, vb);
        }
      }
      if (OP == OP_HPSI) {
        tmem_wait_st();
        tmem_fence_before();
        bar_sync_n(BAR_PAIR + q, 32 * MW);
        tmem_fence_after();
        // pass 3: for each residue class b the 16-point transform over k1 -> the kept rows y = (7a + 16b) mod 112
#define QB200_Y3(B) if (B >= blo && B < bhi) y_pass3_gt<SH, B>(A0 + (i & 1) * ABUF + xc, t0, act);
        QB200_Y3(0) QB200_Y3(1) QB200_Y3(2) QB200_Y3(3) QB200_Y3(4) QB200_Y3(5) QB200_Y3(6)
#undef QB200_Y3
        __threadfence_block();
        bar_arrive_n(BAR_DONE + (i & 1), NT);      // the X warps may take the buffer for the way back
      }
    }
  } else {
    // ------------------------------------------------------------------------------------------ X warps: rows in shared memory
    const int xt = tid - NYT;
    auto xsync = []() { bar_sync_n(BAR_X, NXT); };
    const uint32_t row_bytes = (uint32_t)nvec * 16u;
    uint32_t sphase = 0;
    if (xt == 0 && first < nunits) {
      mbar_expect_tx(&mbar, row_bytes);
      bulk_g2s(stg, zt + ((size_t)first * P.np2 + z) * nvec, row_bytes, &mbar);
    }
    // way back of the unit that used buffer ib: x-DIF, gather (H psi); then clear the in-range positions no column covers
    auto finish = [&](int u, int ib) {
      cplx* B = A0 + ib * ABUF;
      bar_sync_n(BAR_DONE + ib, NT);
      if (OP == OP_HPSI) {
        dif_s<-1, np0, 1, NK, DenseRowsW<pitch>, SH::XSPLIT, SH::XSKIP, false, true, 0, FX.nf - 1>(xt, NXT, B, tw0, xsync);
        xsync();
        cplx* ztrow = zt + ((size_t)u * P.np2 + z) * nvec;
        for (int j = xt; j < nvec; j += NXT) ztrow[j] = B[tpos[j]];
      }
      for (int j = xt; j < nzero; j += NXT) B[tzero[j]] = make_double2(0.0, 0.0);
    };
    int prev = -1, i = 0;
    for (int unit = first; unit < nunits; i++) {
      const int nxt = next_unit(unit);
      cplx* A = A0 + (i & 1) * ABUF;
      mbar_wait(&mbar, sphase);
      sphase ^= 1u;
      for (int j = xt; j < nvec; j += NXT) A[tpos[j]] = stg[j];
      xsync();                                     // scatter complete; every reader of the staging row is done
      if (xt == 0 && nxt < nunits) {
        mbar_expect_tx(&mbar, row_bytes);
        bulk_g2s(stg, zt + ((size_t)nxt * P.np2 + z) * nvec, row_bytes, &mbar);
      }
      // x direction: kept rows, digit-reversed (zeros outside the sphere's h range) -> natural
      dit_s<+1, np0, 1, NK, DenseRowsW<pitch>, SH::XSPLIT, SH::XSKIP, true, false, FX.nf - 1>(xt, NXT, A, tw0, xsync);
      __threadfence_block();
      bar_arrive_n(BAR_FULL + (i & 1), NT);
      if (prev >= 0) finish(prev, (i - 1) & 1);
      prev = unit;
      unit = nxt;
    }
    if (prev >= 0) finish(prev, (i - 1) & 1);
  }
  tmem_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc512(tbase);
}

// ------------------------------------------------------------------------------------------------ density, rho in shared memory
// k_plane_td<SH, NYW, NXW>: the density build of k_plane_t with the CTA's plane of rho_part ACCUMULATED IN SHARED MEMORY over
// all its units (one read-modify-write of global memory per point and CTA instead of one L2 reduction per point and state:
// 1.08 G red.global.add.f64 per MgO216 launch).  The 100 KB plane takes the place of the second buffer of kept rows: the kept
// rows are single-buffered, which costs little here -- the y direction needs them only in its first pass (they are free again
// while the radix-7 pass and the |psi|^2 accumulation run, and that is when the X warps scatter and transform the next unit).
template <class SH> QB200_HD constexpr size_t plane_td_smem(int nvec)
{
  constexpr FftDesc FX = make_fft_desc(SH::NP0);
  return (size_t)((FX.twsize + 7) & ~7) * 16 + (size_t)SH::NKEEP * SH::PITCH * 16 + (size_t)((nvec + 7) & ~7) * 16 + (size_t)SH::NP0 * SH::NP1 * 8;
}

template <class SH, int NYW, int NXW>
__global__ void __launch_bounds__((NYW + NXW) * 32, 1) k_plane_td(const __grid_constant__ DevPlan P, const cplx* __restrict__ zt, double* __restrict__ rho_part,
                                                                  const double* __restrict__ fac, int nunits)
{
  static_assert(SH::NP1 == 112 && NYW % 4 == 0 && NYW >= 4 && NYW <= 16, "thread-per-column y passes are written for 112 = 16 x 7");
  constexpr FftDesc FX = make_fft_desc(SH::NP0);
  constexpr int np0 = SH::NP0, np1 = SH::NP1, pitch = SH::PITCH, np01 = np0 * np1, NK = SH::NKEEP;
  constexpr int NYT = NYW * 32, NXT = NXW * 32, NT = NYT + NXT;
  constexpr int ABUF = NK * pitch;
  extern __shared__ __align__(128) unsigned char smraw[];
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(8) uint64_t mbar;
  const int nvec = P.nvec, nvp = (nvec + 7) & ~7, nzero = P.ntzero;
  cplx* tw0 = reinterpret_cast<cplx*>(smraw);
  cplx* A = tw0 + ((FX.twsize + 7) & ~7);
  cplx* stg = A + ABUF;
  double* acc = reinterpret_cast<double*>(stg + nvp);
  const unsigned short* __restrict__ tpos = P.tpos;
  const unsigned short* __restrict__ tzero = P.tzero;
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
  const int z = blockIdx.x, G = gridDim.y;
  const size_t N = (size_t)np01 * P.np2;
  if (warp == 0) tmem_alloc512(&tmem_slot);
  if (tid == 32) mbar_init(&mbar, 1);
  for (int i = tid; i < FX.twsize; i += NT) tw0[i] = P.tw0p[i];
  for (int i = tid; i < ABUF; i += NT) A[i] = make_double2(0.0, 0.0);
  for (int i = tid; i < np01; i += NT) acc[i] = 0.0;
  tmem_fence_before();
  __syncthreads();
  tmem_fence_after();
  const uint32_t tbase = tmem_slot;
  auto next_unit = [&](int u) {
    u += G;
    while (u < nunits && !fac_active(P, fac, u)) u += G;
    return u;
  };
  const int first = next_unit((int)blockIdx.y - G);
  if (warp < NYW) {
    const int q = warp & 3, m = warp >> 2;
    const uint32_t t0 = tbase + ((uint32_t)(q * 32) << 16);
    const bool act = lane < 28 && 28 * q + lane < np0;
    const int xc = min(28 * q + lane, np0 - 1);
    constexpr int MW = NYW / 4;
    const int blo = (7 * m) / MW, bhi = (7 * (m + 1)) / MW;
    const int klo = (16 * m) / MW, khi = (16 * (m + 1)) / MW;
    constexpr unsigned MASK = zmask(16, 7, SH::YSPLIT, SH::YSKIP);
    const cplx* Ax = A + xc;
    double* ax = acc + xc;
    for (int unit = first; unit < nunits; unit = next_unit(unit)) {
      const double facu = fac_first(fac, unit), facv = fac_second(P, fac, unit);
      bar_sync_n(BAR_FULL, NT);
#pragma unroll 1
      for (int b = blo; b < bhi; b++) {
        cplx x[16];
#pragma unroll
        for (int a = 0; a < 16; a++) {
          const int c = zclass(a, 7, SH::YSPLIT, SH::YSKIP);
          if (c == 0) continue;
          const int y = 7 * a + b;
          const int row = (7 * a + 6 < SH::YSPLIT) ? y : ((7 * a >= SH::YSPLIT + SH::YSKIP) ? y - SH::YSKIP : (y < SH::YSPLIT ? y : y - SH::YSKIP));
          if (c == 1) x[a] = Ax[row * pitch];
          else x[a] = (y < SH::YSPLIT || y >= SH::YSPLIT + SH::YSKIP) ? Ax[row * pitch] : make_double2(0.0, 0.0);
        }
        DftM<16, +1, MASK>::run(x);
        if (b != 0) {
#pragma unroll
          for (int k1 = 1; k1 < 16; k1++) { const double2 w = c_ytw[16 * b + k1]; x[k1] = cmul_s<+1>(x[k1], w.x, w.y); }
        }
        Tmem<16>::st(t0 + 64 * b, x);
      }
      tmem_wait_st();
      tmem_fence_before();
      bar_sync_n(BAR_PAIR + q, 32 * MW);
      tmem_fence_after();
      bar_arrive_n(BAR_DONE, NT);                  // the kept rows are consumed: the X warps may bring the next unit
#pragma unroll 1
      for (int k1 = klo; k1 < khi; k1++) {
        cplx t[7];
        Tmem<1, 7>::ld(t, t0 + 4 * k1, 64);
        Dft<7, +1>::run(t);
        if (act) {                                 // this thread owns (y = k1 + 16 k2, x) in every unit of the CTA
#pragma unroll
          for (int k2 = 0; k2 < 7; k2++) ax[(k1 + 16 * k2) * np0] += facu * t[k2].x * t[k2].x + facv * t[k2].y * t[k2].y;
        }
      }
      // (the other warps of the quarter finish their share of pass 2 before pass 1 of the next unit overwrites the slots)
      tmem_fence_before();
      bar_sync_n(BAR_PAIR + q, 32 * MW);
      tmem_fence_after();
    }
  } else {
    const int xt = tid - NYT;
    auto xsync = []() { bar_sync_n(BAR_X, NXT); };
    const uint32_t row_bytes = (uint32_t)nvec * 16u;
    uint32_t sphase = 0;
    if (xt == 0 && first < nunits) {
      mbar_expect_tx(&mbar, row_bytes);
      bulk_g2s(stg, zt + ((size_t)first * P.np2 + z) * nvec, row_bytes, &mbar);
    }
    int i = 0;
    for (int unit = first; unit < nunits; i++) {
      const int nxt = next_unit(unit);
      mbar_wait(&mbar, sphase);
      sphase ^= 1u;
      if (i > 0) {
        bar_sync_n(BAR_DONE, NT);                  // the Y warps have read the kept rows of the unit before
        for (int j = xt; j < nzero; j += NXT) A[tzero[j]] = make_double2(0.0, 0.0);
      }
      for (int j = xt; j < nvec; j += NXT) A[tpos[j]] = stg[j];
      xsync();
      if (xt == 0 && nxt < nunits) {
        mbar_expect_tx(&mbar, row_bytes);
        bulk_g2s(stg, zt + ((size_t)nxt * P.np2 + z) * nvec, row_bytes, &mbar);
      }
      dit_s<+1, np0, 1, NK, DenseRowsW<pitch>, SH::XSPLIT, SH::XSKIP, true, false, FX.nf - 1>(xt, NXT, A, tw0, xsync);
      __threadfence_block();
      bar_arrive_n(BAR_FULL, NT);
      unit = nxt;
    }
    if (i > 0) bar_sync_n(BAR_DONE, NT);
  }
  tmem_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc512(tbase);
  double* rz = rho_part + (size_t)blockIdx.y * N + (size_t)z * np01;
  for (int i = tid; i < np01; i += NT) rz[i] += acc[i];
}

}  // namespace qb200
