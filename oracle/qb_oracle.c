/* oracle/qb_oracle.c -- TEST INFRASTRUCTURE ONLY (see qb_oracle.h).
 *
 * Plain-C restatement of the reference's per-state H psi / density path.  Each function cites the reference
 * file:line it follows.  Parity is PINNED against the compiled reference (oracle/_ref) and tests/golden/.
 * The 1-D FFT here is an independent textbook mixed-radix decimation-in-time DFT (the reference delegates to
 * FFTW/ESSL or its built-in cfftm; only the transform's definition matters: sign, scaling, pruning).
 */
#include "qb_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

/* ------------------------------------------------------------------------------------------------ small vectors */
typedef struct { double x, y, z; } v3;
static v3 v3s(double a, v3 b) { v3 r = { b.x * a, b.y * a, b.z * a }; return r; }          /* a*b (D3vector *=)  */
static v3 v3add(v3 a, v3 b) { v3 r = { a.x + b.x, a.y + b.y, a.z + b.z }; return r; }
static v3 v3cross(v3 a, v3 b) { v3 r = { a.y*b.z - a.z*b.y, a.z*b.x - a.x*b.z, a.x*b.y - a.y*b.x }; return r; }
static double v3dot(v3 a, v3 b) { return a.x*b.x + a.y*b.y + a.z*b.z; }
static double v3norm(v3 a) { return a.x*a.x + a.y*a.y + a.z*a.z; }                         /* math/d3vector.h:163 */
static double v3len(v3 a) { return sqrt(a.x*a.x + a.y*a.y + a.z*a.z); }

/* UnitCell::set (UnitCell.cc:37-94): volume, b_i = 2 pi (fac*a_j) ^ a_k */
static double cell_recip(const double cell[9], v3 a[3], v3 b[3])
{
  for (int i = 0; i < 3; i++) { a[i].x = cell[3*i]; a[i].y = cell[3*i+1]; a[i].z = cell[3*i+2]; }
  double vol = v3dot(a[0], v3cross(a[1], a[2]));
  double fac = 1.0 / vol;
  b[0] = v3s(2.0 * M_PI, v3cross(v3s(fac, a[1]), a[2]));
  b[1] = v3s(2.0 * M_PI, v3cross(v3s(fac, a[2]), a[0]));
  b[2] = v3s(2.0 * M_PI, v3cross(v3s(fac, a[0]), a[1]));
  return vol;
}

/* ------------------------------------------------------------------------------------------------ Basis */
int qbo_factorizable(int n)   /* Basis.cc:126-147 */
{
  if (n % 11 == 0) n /= 11;
  if (n % 7 == 0) n /= 7;
  if (n % 5 == 0) n /= 5;
  if (n % 3 == 0) n /= 3;
  if (n % 3 == 0) n /= 3;
  while (n % 2 == 0) n /= 2;
  return n == 1;
}

typedef struct { int h, k, lmin, size, seq; } rod_t;
static int rod_cmp(const void* pa, const void* pb)
{
  /* multiset<Rod> ordered by size (Basis.cc:190-193); equal sizes keep insertion order */
  const rod_t* a = (const rod_t*)pa; const rod_t* b = (const rod_t*)pb;
  if (a->size != b->size) return a->size < b->size ? -1 : 1;
  return a->seq < b->seq ? -1 : (a->seq > b->seq);
}

static int imin(int a, int b) { return a < b ? a : b; }
static int imax(int a, int b) { return a > b ? a : b; }

qbo_basis* qbo_basis_create(const double cell[9], double ecut, const double kpoint[3], int force_complex)
{
  qbo_basis* B = (qbo_basis*)calloc(1, sizeof(qbo_basis));
  v3 a[3], b[3];
  B->omega = cell_recip(cell, a, b);
  for (int i = 0; i < 3; i++) { B->b[3*i] = b[i].x; B->b[3*i+1] = b[i].y; B->b[3*i+2] = b[i].z; }
  B->is_real = (kpoint[0] == 0.0 && kpoint[1] == 0.0 && kpoint[2] == 0.0 && !force_complex);  /* Basis.cc:284 */
  const double two_ecut = 2.0 * ecut, twopi = 2.0 * M_PI;
  const double kpx = kpoint[0], kpy = kpoint[1], kpz = kpoint[2];
  const double b2inv2 = 1.0 / v3norm(b[2]);
  const double fac = sqrt(two_ecut) / twopi;
  const int hmax = (int)(0.5 + fac * v3len(a[0])), hmin = -hmax;           /* Basis.cc:397-404 */
  const int kmax = (int)(0.5 + fac * v3len(a[1])), kmin = -kmax;
  const int lmax = (int)(0.5 + fac * v3len(a[2])), lmin = -lmax;
  int cap = (2*hmax + 4) * (2*kmax + 4) + 8, nr = 0;
  rod_t* rods = (rod_t*)malloc(sizeof(rod_t) * cap);
  int hmax_u = hmin, hmin_u = hmax, kmax_u = kmin, kmin_u = kmax, lmax_u = lmin, lmin_u = lmax;
  if (B->is_real) {                                                        /* Basis.cc:417-496 */
    int lend0 = (int)(sqrt(two_ecut * b2inv2));
    rod_t r0 = { 0, 0, 0, lend0 + 1, nr }; rods[nr++] = r0;
    hmax_u = hmin_u = kmin_u = kmax_u = lmin_u = 0; lmax_u = lend0;
    for (int k = 1; k <= kmax + 1; k++) {
      int lstart = lmax, lend = lmin, found = 0;
      for (int l = lmin - 1; l <= lmax + 1; l++) {
        double two_e = v3norm(v3add(v3s(k, b[1]), v3s(l, b[2])));
        if (two_e < two_ecut) { lstart = imin(l, lstart); lend = imax(l, lend); found = 1; }
      }
      if (found) {
        rod_t r = { 0, k, lstart, lend - lstart + 1, nr }; rods[nr++] = r;
        kmax_u = imax(k, kmax_u); kmin_u = imin(k, kmin_u); lmax_u = imax(lend, lmax_u); lmin_u = imin(lstart, lmin_u);
      }
    }
    for (int h = 1; h <= hmax + 1; h++)
      for (int k = kmin - 1; k <= kmax + 1; k++) {
        int lstart = lmax, lend = lmin, found = 0;
        for (int l = lmin - 1; l <= lmax + 1; l++) {
          double two_e = v3norm(v3add(v3add(v3s(h, b[0]), v3s(k, b[1])), v3s(l, b[2])));
          if (two_e < two_ecut) { lstart = imin(l, lstart); lend = imax(l, lend); found = 1; }
        }
        if (found) {
          rod_t r = { h, k, lstart, lend - lstart + 1, nr }; rods[nr++] = r;
          hmax_u = imax(h, hmax_u); hmin_u = imin(h, hmin_u); kmax_u = imax(k, kmax_u); kmin_u = imin(k, kmin_u);
          lmax_u = imax(lend, lmax_u); lmin_u = imin(lstart, lmin_u);
        }
      }
  } else {                                                                 /* Basis.cc:497-536 */
    for (int h = hmin - 1; h <= hmax + 1; h++)
      for (int k = kmin - 1; k <= kmax + 1; k++) {
        int lstart = lmax, lend = lmin, found = 0;
        for (int l = lmin - 1; l <= lmax + 1; l++) {
          double two_e = v3norm(v3add(v3add(v3s(kpx + h, b[0]), v3s(kpy + k, b[1])), v3s(kpz + l, b[2])));
          if (two_e < two_ecut) { lstart = imin(l, lstart); lend = imax(l, lend); found = 1; }
        }
        if (found) {
          rod_t r = { h, k, lstart, lend - lstart + 1, nr }; rods[nr++] = r;
          hmax_u = imax(h, hmax_u); hmin_u = imin(h, hmin_u); kmax_u = imax(k, kmax_u); kmin_u = imin(k, kmin_u);
          lmax_u = imax(lend, lmax_u); lmin_u = imin(lstart, lmin_u);
        }
      }
  }
  B->idxmax[0] = hmax_u; B->idxmin[0] = hmin_u; B->idxmax[1] = kmax_u; B->idxmin[1] = kmin_u;
  B->idxmax[2] = lmax_u; B->idxmin[2] = lmin_u;
  int n;                                                                   /* Basis.cc:563-574 */
  n = 2*hmax + 2; while (!qbo_factorizable(n)) n += 2; B->np[0] = n;
  n = 2*kmax + 2; while (!qbo_factorizable(n)) n += 2; B->np[1] = n;
  n = 2*lmax + 2; while (!qbo_factorizable(n)) n += 2; B->np[2] = n;

  /* one process row: rods in multiset order, then rod(0,0) swapped into slot 0 (Basis.cc:595-651) */
  qsort(rods, nr, sizeof(rod_t), rod_cmp);
  int rank0 = -1;
  for (int i = 0; i < nr; i++) if (rods[i].h == 0 && rods[i].k == 0) rank0 = i;
  if (rank0 > 0) { rod_t t = rods[0]; rods[0] = rods[rank0]; rods[rank0] = t; }
  B->nrods = nr;
  B->rod_h = (int*)malloc(sizeof(int) * nr); B->rod_k = (int*)malloc(sizeof(int) * nr);
  B->rod_lmin = (int*)malloc(sizeof(int) * nr); B->rod_size = (int*)malloc(sizeof(int) * nr);
  B->rod_first = (int*)malloc(sizeof(int) * nr);
  int ngw = 0;
  for (int i = 0; i < nr; i++) {
    B->rod_h[i] = rods[i].h; B->rod_k[i] = rods[i].k; B->rod_lmin[i] = rods[i].lmin; B->rod_size[i] = rods[i].size;
    B->rod_first[i] = ngw; ngw += rods[i].size;
  }
  free(rods);
  B->ngw = ngw;
  B->idx = (int*)malloc(sizeof(int) * 3 * (size_t)ngw);
  B->kpg2 = (double*)malloc(sizeof(double) * (size_t)ngw);
  B->kpgx = (double*)malloc(sizeof(double) * 3 * (size_t)ngw);
  int i = 0;
  for (int irod = 0; irod < nr; irod++)
    for (int l = 0; l < B->rod_size[irod]; l++, i++) {
      B->idx[3*i] = B->rod_h[irod]; B->idx[3*i+1] = B->rod_k[irod]; B->idx[3*i+2] = B->rod_lmin[irod] + l;
    }
  for (i = 0; i < ngw; i++) {                                              /* Basis::update_g, Basis.cc:703-741 */
    v3 kpgt = v3add(v3add(v3s(kpx + B->idx[3*i], b[0]), v3s(kpy + B->idx[3*i+1], b[1])), v3s(kpz + B->idx[3*i+2], b[2]));
    B->kpgx[i] = kpgt.x; B->kpgx[ngw + i] = kpgt.y; B->kpgx[2*(size_t)ngw + i] = kpgt.z;
    B->kpg2[i] = v3norm(kpgt);
  }
  return B;
}

void qbo_basis_destroy(qbo_basis* B)
{
  if (!B) return;
  free(B->rod_h); free(B->rod_k); free(B->rod_lmin); free(B->rod_size); free(B->rod_first);
  free(B->idx); free(B->kpg2); free(B->kpgx); free(B);
}

void qbo_density_grid(const double cell[9], double ecut, int grid[3])   /* ChargeDensity.cc:77-99 */
{
  /* only np of the Gamma basis at 4*ecut is needed, which depends on hmax/kmax/lmax alone (Basis.cc:394-404,563-574) */
  v3 a[3], b[3];
  cell_recip(cell, a, b);
  const double fac = sqrt(2.0 * 4.0 * ecut) / (2.0 * M_PI);
  for (int d = 0; d < 3; d++) {
    int m = (int)(0.5 + fac * v3len(a[d]));
    int n = 2*m + 2; while (!qbo_factorizable(n)) n += 2;
    n += 2; while (!qbo_factorizable(n)) n += 2;
    grid[d] = n;
  }
}

/* ------------------------------------------------------------------------------------------------ 1-D FFT */
typedef struct { int n; double* w; /* w[2k],w[2k+1] = cos,sin(2 pi k/n) */ } fftplan;

static void fftplan_init(fftplan* p, int n)
{
  p->n = n; p->w = (double*)malloc(sizeof(double) * 2 * (size_t)n);
  for (int k = 0; k < n; k++) {
    long double ang = 2.0L * 3.14159265358979323846264338327950288L * (long double)k / (long double)n;
    p->w[2*k] = (double)cosl(ang); p->w[2*k+1] = (double)sinl(ang);
  }
}
static int smallest_factor(int n) { for (int p = 2; p * p <= n; p++) if (n % p == 0) return p; return n; }

/* out[k] = sum_j in[j*is] * exp(sign * 2 pi i jk / n), recursive mixed radix; N = plan length, n divides N */
static void fft_rec(const fftplan* P, int n, int sign, const double* in, int is, double* out)
{
  if (n == 1) { out[0] = in[0]; out[1] = in[1]; return; }
  const int p = smallest_factor(n), m = n / p, N = P->n, tstride = N / n;
  for (int r = 0; r < p; r++) fft_rec(P, m, sign, in + 2 * (size_t)r * is, is * p, out + 2 * (size_t)r * m);
  double tr[16], ti[16];
  double *tre = tr, *tim = ti;
  if (p > 16) { tre = (double*)malloc(sizeof(double) * 2 * p); tim = tre + p; }
  for (int k = 0; k < m; k++) {
    for (int r = 0; r < p; r++) { tre[r] = out[2*(r*m + k)]; tim[r] = out[2*(r*m + k) + 1]; }
    for (int q = 0; q < p; q++) {
      const int kk = k + q * m;
      double sr = 0.0, si = 0.0;
      for (int r = 0; r < p; r++) {
        const int e = (int)(((long long)r * kk) % n) * tstride;
        const double wr = P->w[2*e], wi = sign > 0 ? P->w[2*e+1] : -P->w[2*e+1];
        sr += tre[r] * wr - tim[r] * wi;
        si += tre[r] * wi + tim[r] * wr;
      }
      out[2*kk] = sr; out[2*kk+1] = si;
    }
  }
  if (p > 16) free(tre);
}

/* in-place strided transform of one line: data[j*stride], j<n */
static void fft_line(const fftplan* P, int sign, double* data, int stride, double* tmp)
{
  fft_rec(P, P->n, sign, data, stride, tmp);
  for (int j = 0; j < P->n; j++) { data[2*(size_t)j*stride] = tmp[2*j]; data[2*(size_t)j*stride+1] = tmp[2*j+1]; }
}

/* ------------------------------------------------------------------------------------------------ FourierTransform */
struct qbo_ft {
  int np0, np1, np2, nvec, ntrans0, ngw, is_real;
  int *ifftp, *ifftm;      /* sphere -> zvec index (FourierTransform.cc:262-331, 437-454) */
  int *colxy;              /* column ivec -> hp + np0*kp (the l=0 entry of iunpack_, :361-430, :484-506) */
  double* zvec;            /* nvec*np2 complex */
  fftplan P0, P1, P2;
};

qbo_ft* qbo_ft_create(int np0, int np1, int np2, int nrods, const int* rod_h, const int* rod_k, const int* rod_lmin,
                      const int* rod_size, int is_real, int idxmin1, int idxmax1)
{
  qbo_ft* ft = (qbo_ft*)calloc(1, sizeof(qbo_ft));
  ft->np0 = np0; ft->np1 = np1; ft->np2 = np2; ft->is_real = is_real;
  ft->nvec = is_real ? 2 * nrods - 1 : nrods;                                        /* :186-197 */
  ft->ntrans0 = imax(abs(idxmax1), abs(idxmin1)) + 1;                                /* :202 */
  int ngw = 0; for (int i = 0; i < nrods; i++) ngw += rod_size[i];
  ft->ngw = ngw;
  ft->ifftp = (int*)malloc(sizeof(int) * (size_t)(ngw > 0 ? ngw : 1));
  ft->ifftm = (int*)malloc(sizeof(int) * (size_t)(ngw > 0 ? ngw : 1));
  ft->colxy = (int*)malloc(sizeof(int) * (size_t)(ft->nvec > 0 ? ft->nvec : 1));
  ft->zvec = (double*)malloc(sizeof(double) * 2 * (size_t)ft->nvec * np2 + 16);
  int ig = 0;
  if (is_real) {
    ft->ifftp[0] = 0; ft->ifftm[0] = 0; ig = 1;                                      /* :277-286 */
    for (int l = 1; l < rod_size[0]; l++, ig++) { ft->ifftp[ig] = l; ft->ifftm[ig] = np2 - l; }
    ft->colxy[0] = 0;
    for (int irod = 1; irod < nrods; irod++) {                                       /* :291-305, :372-398 */
      for (int i = 0; i < rod_size[irod]; i++, ig++) {
        const int l = i + rod_lmin[irod];
        int izp = l, izm = -l;
        if (izp < 0) izp += np2;
        if (izm < 0) izm += np2;
        ft->ifftp[ig] = (2*irod - 1) * np2 + izp;
        ft->ifftm[ig] = (2*irod) * np2 + izm;
      }
      int hp = rod_h[irod], kp = rod_k[irod];
      if (hp < 0) hp += np0;
      if (kp < 0) kp += np1;
      int hm = -hp, km = -kp;
      if (hm < 0) hm += np0;
      if (km < 0) km += np1;
      ft->colxy[2*irod - 1] = hp + np0 * kp;
      ft->colxy[2*irod] = hm + np0 * km;
    }
  } else {
    for (int irod = 0; irod < nrods; irod++) {                                       /* :441-453, :486-506 */
      for (int i = 0; i < rod_size[irod]; i++, ig++) {
        int iz = i + rod_lmin[irod];
        if (iz < 0) iz += np2;
        ft->ifftp[ig] = irod * np2 + iz;
      }
      int h = rod_h[irod], k = rod_k[irod];
      if (h < 0) h += np0;
      if (k < 0) k += np1;
      ft->colxy[irod] = h + np0 * k;
    }
  }
  fftplan_init(&ft->P0, np0); fftplan_init(&ft->P1, np1); fftplan_init(&ft->P2, np2);
  return ft;
}

void qbo_ft_destroy(qbo_ft* ft)
{
  if (!ft) return;
  free(ft->ifftp); free(ft->ifftm); free(ft->colxy); free(ft->zvec); free(ft->P0.w); free(ft->P1.w); free(ft->P2.w); free(ft);
}
int qbo_ft_nvec(const qbo_ft* ft) { return ft->nvec; }
int qbo_ft_ntrans0(const qbo_ft* ft) { return ft->ntrans0; }

/* FourierTransform::bwd (FourierTransform.cc:584-980), one rank: z-FFT (+1), zero val, scatter columns, x-FFTs on the
 * rows j in [0,ntrans0) U [np1-ntrans0,np1), y-FFTs on every column; no scaling. */
static void bwd(qbo_ft* ft, double* val)
{
  const int np0 = ft->np0, np1 = ft->np1, np2 = ft->np2, nvec = ft->nvec;
  const size_t np01 = (size_t)np0 * np1;
  #pragma omp parallel
  {
    double* tmp = (double*)malloc(sizeof(double) * 2 * (size_t)imax(np0, imax(np1, np2)));
    #pragma omp for
    for (int iv = 0; iv < nvec; iv++) fft_line(&ft->P2, +1, ft->zvec + 2 * (size_t)iv * np2, 1, tmp);
    #pragma omp for
    for (int k = 0; k < np2; k++) {
      double* pl = val + 2 * np01 * k;
      memset(pl, 0, sizeof(double) * 2 * np01);
      for (int iv = 0; iv < nvec; iv++) {
        pl[2*ft->colxy[iv]] = ft->zvec[2*((size_t)iv*np2 + k)];
        pl[2*ft->colxy[iv]+1] = ft->zvec[2*((size_t)iv*np2 + k)+1];
      }
      int nlo = ft->ntrans0 < np1 ? ft->ntrans0 : np1;
      for (int j = 0; j < np1; j++)
        if (j < nlo || j >= np1 - ft->ntrans0) fft_line(&ft->P0, +1, pl + 2 * (size_t)j * np0, 1, tmp);
      for (int i = 0; i < np0; i++) fft_line(&ft->P1, +1, pl + 2 * i, np0, tmp);
    }
    free(tmp);
  }
}

/* FourierTransform::fwd (FourierTransform.cc:983-1361): y-FFTs (-1), x-FFTs on the kept rows, gather columns,
 * z-FFT (-1), scale 1/(np0*np1*np2) (:1338-1342). val is clobbered, as in the reference. */
static void fwd(qbo_ft* ft, double* val)
{
  const int np0 = ft->np0, np1 = ft->np1, np2 = ft->np2, nvec = ft->nvec;
  const size_t np01 = (size_t)np0 * np1;
  const double fac = 1.0 / ((double)np0 * np1 * np2);
  #pragma omp parallel
  {
    double* tmp = (double*)malloc(sizeof(double) * 2 * (size_t)imax(np0, imax(np1, np2)));
    #pragma omp for
    for (int k = 0; k < np2; k++) {
      double* pl = val + 2 * np01 * k;
      for (int i = 0; i < np0; i++) fft_line(&ft->P1, -1, pl + 2 * i, np0, tmp);
      int nlo = ft->ntrans0 < np1 ? ft->ntrans0 : np1;
      for (int j = 0; j < np1; j++)
        if (j < nlo || j >= np1 - ft->ntrans0) fft_line(&ft->P0, -1, pl + 2 * (size_t)j * np0, 1, tmp);
      for (int iv = 0; iv < nvec; iv++) {
        ft->zvec[2*((size_t)iv*np2 + k)] = pl[2*ft->colxy[iv]];
        ft->zvec[2*((size_t)iv*np2 + k)+1] = pl[2*ft->colxy[iv]+1];
      }
    }
    #pragma omp for
    for (int iv = 0; iv < nvec; iv++) {
      double* col = ft->zvec + 2 * (size_t)iv * np2;
      fft_line(&ft->P2, -1, col, 1, tmp);
      for (int k = 0; k < 2 * np2; k++) col[k] *= fac;
    }
    free(tmp);
  }
}

void qbo_backward(qbo_ft* ft, const double* c, double* f)
{
  /* vector_to_zvec (FourierTransform.cc:1624-1664) */
  double* pz = ft->zvec;
  memset(pz, 0, sizeof(double) * 2 * (size_t)ft->nvec * ft->np2);
  for (int ig = 0; ig < ft->ngw; ig++) {
    const double a = c[2*ig], b = c[2*ig+1];
    const int ip = ft->ifftp[ig];
    pz[2*ip] = a; pz[2*ip+1] = b;
    if (ft->is_real) { const int im = ft->ifftm[ig]; pz[2*im] = a; pz[2*im+1] = -b; }
  }
  bwd(ft, f);
}

void qbo_forward(qbo_ft* ft, double* f, double* c)
{
  fwd(ft, f);
  for (int ig = 0; ig < ft->ngw; ig++) {            /* zvec_to_vector (:1666-1681) */
    const int ip = ft->ifftp[ig];
    c[2*ig] = ft->zvec[2*ip]; c[2*ig+1] = ft->zvec[2*ip+1];
  }
}

void qbo_backward_pair(qbo_ft* ft, const double* c1, const double* c2, double* f)
{
  /* doublevector_to_zvec (:1684-1720) */
  double* pz = ft->zvec;
  memset(pz, 0, sizeof(double) * 2 * (size_t)ft->nvec * ft->np2);
  for (int ig = 0; ig < ft->ngw; ig++) {
    const double a = c1[2*ig], b = c1[2*ig+1], c = c2[2*ig], d = c2[2*ig+1];
    const int ip = ft->ifftp[ig], im = ft->ifftm[ig];
    pz[2*ip] = a - d; pz[2*ip+1] = b + c;
    pz[2*im] = a + d; pz[2*im+1] = c - b;
  }
  bwd(ft, f);
}

void qbo_forward_pair(qbo_ft* ft, double* f, double* c1, double* c2)
{
  fwd(ft, f);
  const double* pz = ft->zvec;                       /* zvec_to_doublevector (:1723-1752) */
  for (int ig = 0; ig < ft->ngw; ig++) {
    const int ip = ft->ifftp[ig], im = ft->ifftm[ig];
    const double a = pz[2*ip], b = pz[2*ip+1], c = pz[2*im], d = pz[2*im+1];
    c1[2*ig] = 0.5 * (a + c); c1[2*ig+1] = 0.5 * (b - d);
    c2[2*ig] = 0.5 * (b + d); c2[2*ig+1] = 0.5 * (c - a);
  }
}

/* ------------------------------------------------------------------------------------------------ SlaterDet */
void qbo_rs_mul_add(qbo_ft* ft, int ngw, int ldc, int nst, const double* c, const double* v, double* cp)
{
  /* SlaterDet.cc:971-1040.  ctmp has mloc entries of which only ngw are written by forward(); the reference's
   * daxpy/zaxpy over mloc adds ctmp's zero-initialised padding, i.e. nothing. */
  const size_t N = (size_t)ft->np0 * ft->np1 * ft->np2;
  double* tmp = (double*)malloc(sizeof(double) * 2 * N);
  double* ct = (double*)calloc(4 * (size_t)ldc, sizeof(double));
  if (ft->is_real) {
    int n;
    for (n = 0; n < nst - 1; n += 2) {
      qbo_backward_pair(ft, c + 2*(size_t)n*ldc, c + 2*(size_t)(n+1)*ldc, tmp);
      for (size_t i = 0; i < N; i++) { const double vi = v[i]; tmp[2*i] = vi * tmp[2*i]; tmp[2*i+1] = vi * tmp[2*i+1]; }
      qbo_forward_pair(ft, tmp, ct, ct + 2*(size_t)ldc);
      for (int ig = 0; ig < 2*ngw; ig++) { cp[2*(size_t)n*ldc + ig] += ct[ig]; cp[2*(size_t)(n+1)*ldc + ig] += ct[2*(size_t)ldc + ig]; }
    }
    if (nst % 2 != 0) {
      n = nst - 1;
      qbo_backward(ft, c + 2*(size_t)n*ldc, tmp);
      for (size_t i = 0; i < N; i++) { tmp[2*i] = v[i] * tmp[2*i]; tmp[2*i+1] = 0.0; }
      qbo_forward(ft, tmp, ct);
      for (int ig = 0; ig < 2*ngw; ig++) cp[2*(size_t)n*ldc + ig] += ct[ig];
    }
  } else {
    for (int n = 0; n < nst; n++) {
      qbo_backward(ft, c + 2*(size_t)n*ldc, tmp);
      for (size_t i = 0; i < N; i++) { const double vi = v[i]; tmp[2*i] *= vi; tmp[2*i+1] *= vi; }
      qbo_forward(ft, tmp, ct);
      for (int ig = 0; ig < 2*ngw; ig++) cp[2*(size_t)n*ldc + ig] += ct[ig];
    }
  }
  free(tmp); free(ct);
}

void qbo_compute_density(qbo_ft* ft, int ldc, int nst, const double* c, const double* fac, double* rho)
{
  /* SlaterDet.cc:905-926 (the pair branch is disabled in the reference, :858) */
  const size_t N = (size_t)ft->np0 * ft->np1 * ft->np2;
  double* tmp = (double*)malloc(sizeof(double) * 2 * N);
  for (int n = 0; n < nst; n++) {
    if (fac[n] > 0.0) {
      qbo_backward(ft, c + 2*(size_t)n*ldc, tmp);
      const double f = fac[n];
      for (size_t i = 0; i < N; i++) rho[i] += f * (tmp[2*i]*tmp[2*i] + tmp[2*i+1]*tmp[2*i+1]);
    }
  }
  free(tmp);
}

void qbo_compute_current(qbo_ft* ft, int ngw, int ldc, int nst, const double* c, const double* fac, const double* kpgx, double* cur)
{
  /* CurrentDensity::update_current (CurrentDensity.cc:52-86) around SlaterDet::compute_density(ft, weight, complex* rho, sd2)
   * (SlaterDet.cc:935-968): per direction rwf = i*kpgx[idir]*c (:72-76), tmp = sum_n fac_n conj(psi_n(r)) rwf_n(r) over
   * the states with fac_n > 0, current[idir][r] += -Im tmp (:85-87).  cur: 3*N doubles, accumulated. */
  const size_t N = (size_t)ft->np0 * ft->np1 * ft->np2;
  double* t1 = (double*)malloc(sizeof(double) * 2 * N);
  double* t2 = (double*)malloc(sizeof(double) * 2 * N);
  double* acc = (double*)malloc(sizeof(double) * 2 * N);
  double* rw = (double*)malloc(sizeof(double) * 2 * (size_t)ngw);
  for (int idir = 0; idir < 3; idir++) {
    const double* k = kpgx + (size_t)idir * ngw;
    memset(acc, 0, sizeof(double) * 2 * N);
    for (int n = 0; n < nst; n++) {
      if (!(fac[n] > 0.0)) continue;
      const double* cn = c + 2 * (size_t)n * ldc;
      for (int ig = 0; ig < ngw; ig++) { rw[2*ig] = -k[ig] * cn[2*ig+1]; rw[2*ig+1] = k[ig] * cn[2*ig]; }   /* (0,1)*k*c */
      qbo_backward(ft, cn, t1);
      qbo_backward(ft, rw, t2);
      const double f = fac[n];
      for (size_t i = 0; i < N; i++) {               /* rho[i] += fac*conj(tmp1[i])*tmp2[i] */
        acc[2*i] += f * (t1[2*i] * t2[2*i] + t1[2*i+1] * t2[2*i+1]);
        acc[2*i+1] += f * (t1[2*i] * t2[2*i+1] - t1[2*i+1] * t2[2*i]);
      }
    }
    for (size_t i = 0; i < N; i++) cur[(size_t)idir * N + i] += -acc[2*i+1];
  }
  free(t1); free(t2); free(acc); free(rw);
}

void qbo_kinetic_add(int ngw, int ldc, int nst, const double* kpg2, const double* c, double* cp)
{
  for (int n = 0; n < nst; n++)                       /* EnergyFunctional.cc:1675-1677 */
    for (int ig = 0; ig < ngw; ig++) {
      const double h = 0.5 * kpg2[ig];
      cp[2*((size_t)n*ldc + ig)] += h * c[2*((size_t)n*ldc + ig)];
      cp[2*((size_t)n*ldc + ig)+1] += h * c[2*((size_t)n*ldc + ig)+1];
    }
}

/* kinetic-energy section of EnergyFunctional::energy (EnergyFunctional.cc:1155-1296) for one (spin, k-point):
 * psi2sum[ig] = fac * sum_n occ[n] |c[ig,n]|^2 (:1209-1223), then the 14 partial sums tsum[] of :1225-1276 (the stress
 * sums only if kpgx != NULL, the confinement sums only if fstress / dfstress != NULL).  Same loop order as the reference. */
void qbo_ekin_sums(int ngw, int ldc, int nst, const double* c, const double* occ, double fac, const double* kpg2,
                   const double* kpgx, const double* fstress, const double* dfstress, double* psi2sum, double* tsum)
{
  for (int ig = 0; ig < ngw; ig++) psi2sum[ig] = 0.0;
  for (int n = 0; n < nst; n++)
    for (int ig = 0; ig < ngw; ig++) {
      const double re = c[2*((size_t)n*ldc + ig)], im = c[2*((size_t)n*ldc + ig)+1];
      psi2sum[ig] += fac * occ[n] * (re*re + im*im);                                     /* :1219-1221 */
    }
  for (int k = 0; k < 14; k++) tsum[k] = 0.0;
  for (int ig = 0; ig < ngw; ig++) {
    const double p2 = psi2sum[ig];
    tsum[0] += p2 * kpg2[ig];                                                            /* :1230 */
    if (kpgx) {
      const double x = kpgx[ig], y = kpgx[ngw + ig], z = kpgx[2*(size_t)ngw + ig], f = 2.0 * p2;   /* :1232-1247 */
      tsum[1] += f*x*x; tsum[2] += f*y*y; tsum[3] += f*z*z; tsum[4] += f*x*y; tsum[5] += f*y*z; tsum[6] += f*x*z;
    }
  }
  if (fstress)
    for (int ig = 0; ig < ngw; ig++) {
      const double p2 = psi2sum[ig];
      tsum[7] += p2 * fstress[ig];                                                       /* :1258 */
      if (kpgx && dfstress) {
        const double x = kpgx[ig], y = kpgx[ngw + ig], z = kpgx[2*(size_t)ngw + ig], f = p2 * dfstress[ig];  /* :1260-1273 */
        tsum[8] += f*x*x; tsum[9] += f*y*y; tsum[10] += f*z*z; tsum[11] += f*x*y; tsum[12] += f*y*z; tsum[13] += f*x*z;
      }
    }
}

/* PSDAWavefunctionStepper::update after the descent direction (PSDAWavefunctionStepper.cc:93-225 real, :281-395 complex)
 * with Preconditioner::apply(sd, ispin, ikp, -1.0) (Preconditioner.cc:118-139), one rank.  Returns theta before clipping. */
double qbo_psda_update(int ngw, int ldc, int nst, int is_real, double* c, double* dc, double* c_last, double* dc_last,
                       const double* occ, const double* precdiag, int extrapolate)
{
  for (int n = 0; n < nst; n++)                                                    /* Preconditioner.cc:127-137 */
    for (int i = 0; i < ngw; i++) {
      dc[2*((size_t)n*ldc + i)]   *= -1.0 * precdiag[i];
      dc[2*((size_t)n*ldc + i)+1] *= -1.0 * precdiag[i];
    }
  const size_t n2 = 2 * (size_t)ldc * nst;
  double theta_raw = 0.0;
  if (extrapolate) {
    double a = 0.0, b = 0.0;
    for (int n = 0; n < nst; n++)
      for (size_t i = 0; i < 2 * (size_t)ldc; i++) {                                /* :124-136 / :334-341 */
        const double f = dc[i + 2*(size_t)ldc*n], df = f - dc_last[i + 2*(size_t)ldc*n];
        a += occ[n] * f * df;
        b += occ[n] * df * df;
      }
    if (is_real) {                                                                  /* :146-173 */
      a *= 2.0; b *= 2.0;
      for (int n = 0; n < nst; n++) {
        const size_t i = 2 * (size_t)ldc * n;
        const double f0 = dc[i], f1 = dc[i+1], d0 = f0 - dc_last[i], d1 = f1 - dc_last[i+1];
        a -= occ[n] * (f0*d0 + f1*d1);
        b -= occ[n] * (d0*d0 + d1*d1);
      }
    }
    double theta = 0.0;
    if (b != 0.0) theta = -a / b;                                                   /* :182-183 */
    theta_raw = theta;
    if (theta < -1.0) theta = 0.0;                                                  /* :188-191 */
    if (theta > 2.0) theta = 2.0;
    for (size_t i = 0; i < n2; i++) {                                               /* :196-209 */
      const double x = c[i], xbar = x + theta * (x - c_last[i]);
      const double f = dc[i], fbar = f + theta * (f - dc_last[i]);
      c[i] = xbar + fbar; c_last[i] = x; dc_last[i] = f;
    }
  } else {
    for (size_t i = 0; i < n2; i++) {                                               /* :212-221 */
      const double x = c[i], f = dc[i];
      c[i] = x + f; c_last[i] = x; dc_last[i] = f;
    }
  }
  return theta_raw;
}

/* ------------------------------------------------------------------------------------------------ XC functionals, update_vhxc */
/* LDAFunctional::xc_unpolarized (functionals/LDAFunctional.cc:96-161): Perdew-Zunger / Ceperley-Alder, unpolarized */
void qbo_xc_lda(size_t n, const double* rho, double* exc, double* vxc)
{
  const double c1 = 0.6203504908994001, c3 = -0.610887057711;
  const double A = 0.0311, B = -0.048, b1 = 1.0529, b2 = 0.3334, G = -0.1423;
  const double D = G / (1.0 + b1 + b2) - B;                                                   /* :119-120 */
  const double C = -A - D - G * ((b1/2.0 + b2) / ((1.0+b1+b2)*(1.0+b1+b2)));
  for (size_t i = 0; i < n; i++) {
    double ee = 0.0, vv = 0.0;
    const double rh = rho[i];
    if (rh > 0.0) {
      const double ro13 = cbrt(rh), rs = c1 / ro13;
      const double vx = c3 / rs, ex = 0.75 * vx;                                              /* :133-134 */
      double ec, vc;
      if (rs < 1.0) {                                                                         /* :137-144 */
        const double logrs = log(rs);
        ec = A * logrs + B + C * rs * logrs + D * rs;
        vc = A * logrs + (B - A / 3.0) + (2.0/3.0) * C * rs * logrs + ((2.0 * D - C) / 3.0) * rs;
      } else {                                                                                /* :146-153 */
        const double sqrtrs = sqrt(rs), den = 1.0 + b1 * sqrtrs + b2 * rs;
        ec = G / den;
        vc = ec * (1.0 + (7.0/6.0) * b1 * sqrtrs + (4.0/3.0) * b2 * rs) / den;
      }
      ee = ex + ec; vv = vx + vc;
    }
    exc[i] = ee; vxc[i] = vv;
  }
}

static void gcor2(double a, double a1, double b1, double b2, double b3, double b4, double rtrs, double* gg, double* ggrs)
{                                                                                             /* PBEFunctional.cc:482-492 */
  const double q0 = -2.0 * a * (1.0 + a1 * rtrs * rtrs);
  const double q1 = 2.0 * a * rtrs * (b1 + rtrs * (b2 + rtrs * (b3 + rtrs * b4)));
  const double q2 = log(1.0 + 1.0 / q1);
  *gg = q0 * q2;
  const double q3 = a * (b1 / rtrs + 2.0 * b2 + rtrs * (3.0 * b3 + 4.0 * b4 * rtrs));
  *ggrs = -2.0 * a * a1 * q2 - q0 * q3 / (q1 * (1.0 + q1));
}

/* PBEFunctional::excpbe (functionals/PBEFunctional.cc:196-291), unpolarized; grad = |grad rho| */
void qbo_xc_pbe(size_t n, const double* rho_, const double* grad_, double* exc, double* vxc1, double* vxc2)
{
  const double third = 1.0/3.0, third4 = 4.0/3.0;
  const double ax = -0.7385587663820224058, um = 0.2195149727645171, uk = 0.804, ul = um / uk;
  const double pi32third = 3.09366772628014, alpha = 1.91915829267751, seven_sixth = 7.0/6.0, four_over_pi = 1.27323954473516;
  const double gamma = 0.03109069086965489, bet = 0.06672455060314922, delt = bet / gamma;
  for (size_t i = 0; i < n; i++) {
    const double rho = rho_[i], grad = grad_[i];
    exc[i] = 0.0; vxc1[i] = 0.0; vxc2[i] = 0.0;
    if (rho < 1.e-18) continue;                                                               /* :219-221 */
    const double rh13 = pow(rho, third), exunif = ax * rh13, fk = pi32third * rh13;
    const double s = grad / (2.0 * fk * rho), s2 = s * s, p0 = 1.0 + ul * s2, fxpbe = 1.0 + uk - uk / p0;
    const double ex = exunif * fxpbe, fs = 2.0 * uk * ul / (p0 * p0);
    const double vx1 = third4 * exunif * (fxpbe - s2 * fs), vx2 = -exunif * fs / (rho * 4.0 * fk * fk);   /* :246-247 */
    const double rs = alpha / fk, twoks = 2.0 * sqrt(four_over_pi * fk), t = grad / (twoks * rho), rtrs = sqrt(rs);
    double ec, ecrs;
    gcor2(0.0310907, 0.2137, 7.5957, 3.5876, 1.6382, 0.49294, rtrs, &ec, &ecrs);
    const double vc = ec - rs * ecrs * third;
    const double pon = -ec / gamma, b = delt / (exp(pon) - 1.0), b2 = b * b, t2 = t * t, t4 = t2 * t2;
    const double q4 = 1.0 + b * t2, q5 = q4 + b2 * t4, h = gamma * log(1.0 + delt * q4 * t2 / q5);
    const double t6 = t4 * t2, rsthrd = rs * third, fac = delt / b + 1.0, bec = b2 * fac / bet;
    const double q8 = q5 * q5 + delt * q4 * q5 * t2, q9 = 1.0 + 2.0 * b * t2;
    const double hb = -bet * b * t6 * (2.0 + b * t2) / q8, hrs = -rsthrd * hb * bec * ecrs, ht = 2.0 * bet * q9 / q8;
    const double vc1 = vc + h + hrs - t2 * ht * seven_sixth, vc2 = -ht / (rho * twoks * twoks);
    exc[i] = ex + ec + h; vxc1[i] = vx1 + vc1; vxc2[i] = vx2 + vc2;                           /* :288-290 */
  }
}

/* EnergyFunctional::update_vhxc (EnergyFunctional.cc:353-975) with XCPotential::update (XCPotential.cc:104-460): one spin, no
 * NLCC / ESM / enthalpy / TDDFT-split density.  ft: the density-basis transform.  xc: 0 LDA, 1 PBE.  energies: exc, eps, ehart.
 * The reference's own sequence of transforms (GGA: 3 backward, then per direction forward / i G_j / backward). */
void qbo_update_vhxc(qbo_ft* ft, int xc, int ng, int is_real, const double* rhor, const double* rhog, const double* gx,
                     const double* g2i, const double* vion, const double* rhopst, double omega, double* v_r, double* energies)
{
  const size_t N = (size_t)ft->np0 * ft->np1 * ft->np2;
  const double omega_inv = 1.0 / omega, fpi = 4.0 * M_PI;
  double* tmpr = (double*)malloc(2 * N * sizeof(double));
  double* tmp1 = (double*)malloc(2 * (size_t)ng * sizeof(double));
  double* exc = (double*)malloc(N * sizeof(double));
  double* v1 = (double*)malloc(N * sizeof(double));
  for (size_t i = 0; i < N; i++) v_r[i] = 0.0;                                                /* :382-384 */
  double esum = 0.0;
  if (xc == 0) {                                                                              /* XCPotential.cc:133-175 */
    qbo_xc_lda(N, rhor, exc, v1);
    for (size_t i = 0; i < N; i++) { esum += rhor[i] * exc[i]; v_r[i] += v1[i]; }
  } else {
    double* gr = (double*)malloc(3 * N * sizeof(double));
    double* gmod = (double*)malloc(N * sizeof(double));
    double* v2 = (double*)malloc(N * sizeof(double));
    double* vxctmp = (double*)malloc(N * sizeof(double));
    for (int j = 0; j < 3; j++) {                                                             /* :200-214 */
      for (int ig = 0; ig < ng; ig++) {
        const double g = omega_inv * gx[(size_t)j * ng + ig];
        tmp1[2*ig] = -g * rhog[2*ig+1]; tmp1[2*ig+1] = g * rhog[2*ig];
      }
      qbo_backward(ft, tmp1, tmpr);
      for (size_t i = 0; i < N; i++) gr[(size_t)j * N + i] = tmpr[2*i];
    }
    for (size_t i = 0; i < N; i++) gmod[i] = sqrt(gr[i]*gr[i] + gr[N+i]*gr[N+i] + gr[2*N+i]*gr[2*N+i]);   /* PBEFunctional.cc:101-103 */
    qbo_xc_pbe(N, rhor, gmod, exc, v1, v2);
    for (int j = 0; j < 3; j++) {                                                             /* XCPotential.cc:262-285 */
      for (size_t i = 0; i < N; i++) { tmpr[2*i] = gr[(size_t)j * N + i] * v2[i]; tmpr[2*i+1] = 0.0; }
      qbo_forward(ft, tmpr, tmp1);
      for (int ig = 0; ig < ng; ig++) {
        const double g = gx[(size_t)j * ng + ig], re = tmp1[2*ig], im = tmp1[2*ig+1];
        tmp1[2*ig] = -g * im; tmp1[2*ig+1] = g * re;
      }
      qbo_backward(ft, tmp1, tmpr);
      for (size_t i = 0; i < N; i++) vxctmp[i] = (j == 0 ? 0.0 : vxctmp[i]) + tmpr[2*i];
    }
    for (size_t i = 0; i < N; i++) { esum += rhor[i] * exc[i]; v_r[i] += v1[i] + vxctmp[i]; }  /* :397-410 */
    free(gr); free(gmod); free(v2); free(vxctmp);
  }
  energies[0] = esum * omega / (double)N;                                                     /* :170, :452 */
  /* eps (EnergyFunctional.cc:447-465) and ehart, vlocal_g (:487-518) */
  double eps = 0.0, ehsum = 0.0;
  for (int ig = 0; ig < ng; ig++) {
    const double rx = omega_inv * rhog[2*ig], ry = omega_inv * rhog[2*ig+1];
    eps += rx * vion[2*ig] + ry * vion[2*ig+1];
  }
  if (is_real) eps = 2.0 * eps - (omega_inv * rhog[0] * vion[0] + omega_inv * rhog[1] * vion[1]);
  energies[1] = eps * omega;
  for (int ig = 0; ig < ng; ig++) {
    const double tx = omega_inv * rhog[2*ig] + rhopst[2*ig], ty = omega_inv * rhog[2*ig+1] + rhopst[2*ig+1];
    ehsum += (tx*tx + ty*ty) * g2i[ig];
    tmp1[2*ig] = vion[2*ig] + fpi * tx * g2i[ig];
    tmp1[2*ig+1] = vion[2*ig+1] + fpi * ty * g2i[ig];
  }
  energies[2] = (is_real ? 1.0 : 0.5) * omega * fpi * ehsum;
  qbo_backward(ft, tmp1, tmpr);                                                               /* :931-939 */
  for (size_t i = 0; i < N; i++) v_r[i] += tmpr[2*i];
  free(tmpr); free(tmp1); free(exc); free(v1);
}

/* ------------------------------------------------------------------------------------------------ NonLocalPotential */
double qbo_nl_energy_species(int ngw, int ldc, int nst, const double* c, const double* occ, int is_real, int na, int npr,
                             const int* lproj, const double* wt, const double* twnl, const double* tau,
                             const double* kpgx, double omega, int na_block_size, int compute_hpsi, double* cp)
{
  if (npr <= 0 || na <= 0) return 0.0;
  const double omega_inv = 1.0 / omega;
  double enl = 0.0;
  const int na_blocks = na / na_block_size + (na % na_block_size == 0 ? 0 : 1);      /* :1920-1921 */
  double* anl = (double*)malloc(sizeof(double) * 2 * (size_t)npr * na_block_size * ngw);
  double* fnl = (double*)malloc(sizeof(double) * 2 * (size_t)npr * na_block_size * nst);
  for (int ib = 0; ib < na_blocks; ib++) {
    const int iastart = ib * na_block_size;
    const int iaend = (ib + 1) * na_block_size < na ? (ib + 1) * na_block_size : na;
    const int nab = iaend - iastart;
    const int nprnaloc = nab * npr;
    /* anl[ig + (ia + ipr*nab)*ngw] = twnl * (-i)^l * exp(i*kpgr), kpgr = -(k+G).tau  (:1959-2036) */
    #pragma omp parallel for collapse(2)
    for (int ipr = 0; ipr < npr; ipr++)
      for (int ia = 0; ia < nab; ia++) {
        const double* t = twnl + (size_t)ngw * ipr;
        const int l = lproj[ipr];
        const double* tu = tau + 3 * (size_t)(iastart + ia);
        double* a = anl + 2 * (size_t)(ia + ipr * nab) * ngw;
        for (int ig = 0; ig < ngw; ig++) {
          /* dgemm with alpha=-1 over k=3 (:1969): -(x*tx + y*ty + z*tz) */
          const double arg = -(kpgx[ig] * tu[0] + kpgx[ngw + ig] * tu[1] + kpgx[2*(size_t)ngw + ig] * tu[2]);
          const double s = sin(arg), co = cos(arg);
          if (l == 0) { a[2*ig] = t[ig] * co; a[2*ig+1] = t[ig] * s; }
          else if (l == 1) { a[2*ig] = t[ig] * s; a[2*ig+1] = -t[ig] * co; }
          else if (l == 2) { a[2*ig] = -t[ig] * co; a[2*ig+1] = -t[ig] * s; }
          else { a[2*ig] = -t[ig] * s; a[2*ig+1] = t[ig] * co; }
        }
      }
    /* fnl = anl^H c  (complex, :2064)  or  anl^T c over 2*ngw reals (Gamma, :2053) */
    #pragma omp parallel for collapse(2)
    for (int n = 0; n < nst; n++)
      for (int p = 0; p < nprnaloc; p++) {
        const double* a = anl + 2 * (size_t)p * ngw;
        const double* cn = c + 2 * (size_t)n * ldc;
        double* f = fnl + 2 * ((size_t)p + (size_t)n * nprnaloc);
        if (is_real) {
          double s = 0.0;
          for (int ig = 0; ig < 2*ngw; ig++) s += a[ig] * cn[ig];
          s += -0.5 * a[0] * cn[0];                 /* dger G=0 double-count fix (:2078-2080) */
          f[0] = 2.0 * s; f[1] = 0.0;               /* factor 2: G and -G (:2102) */
        } else {
          double sr = 0.0, si = 0.0;                /* conj(a)*c */
          for (int ig = 0; ig < ngw; ig++) {
            sr += a[2*ig] * cn[2*ig] + a[2*ig+1] * cn[2*ig+1];
            si += a[2*ig] * cn[2*ig+1] - a[2*ig+1] * cn[2*ig];
          }
          f[0] = sr; f[1] = si;
        }
      }
    /* enl and fnl <- wt/omega * fnl (:2106-2148) */
    for (int ipr = 0; ipr < npr; ipr++) {
      const double fac = wt[ipr] * omega_inv;
      for (int n = 0; n < nst; n++) {
        const double facn = fac * occ[n];
        for (int ia = 0; ia < nab; ia++) {
          double* f = fnl + 2 * ((size_t)(ia + ipr * nab) + (size_t)n * nprnaloc);
          enl += facn * (f[0]*f[0] + f[1]*f[1]);
          f[0] *= fac; f[1] *= fac;
        }
      }
    }
    if (compute_hpsi) {                              /* cp += anl * fnl (:2150-2171) */
      #pragma omp parallel for
      for (int n = 0; n < nst; n++) {
        double* cpn = cp + 2 * (size_t)n * ldc;
        for (int p = 0; p < nprnaloc; p++) {
          const double* a = anl + 2 * (size_t)p * ngw;
          const double fr = fnl[2*((size_t)p + (size_t)n*nprnaloc)], fi = fnl[2*((size_t)p + (size_t)n*nprnaloc)+1];
          if (is_real) { for (int ig = 0; ig < 2*ngw; ig++) cpn[ig] += a[ig] * fr; }
          else for (int ig = 0; ig < ngw; ig++) {
            cpn[2*ig] += a[2*ig] * fr - a[2*ig+1] * fi;
            cpn[2*ig+1] += a[2*ig] * fi + a[2*ig+1] * fr;
          }
        }
      }
    }
  }
  free(anl); free(fnl);
  return enl;
}

/* ------------------------------------------------------------------------------------------------ subspace dense LA (f1)
 * Storage as ComplexMatrix::val (math/matrix.h:294-340): column-major ldc x nst complex; real bases are read through the
 * DoubleMatrix proxy (2*ldc real rows).  nall = columns of c (all states), hc holds nst columns (the local shard). */
void qbo_residual(int ldc, int nall, int nst, int is_real, const double* c, double* hc, double* a)
{
  /* PSDAWavefunctionStepper.cc:65-84 (real) / :264-277 (complex); PSDWavefunctionStepper.cc:62-90 is identical */
  const size_t m2 = 2 * (size_t)ldc;
  if (is_real) {
    /* a = 2 c^T cp (gemm 't','n', 2.0) ; a -= c(row 0)^T cp(row 0) (ger -1.0) ; cp -= c a (gemm 'n','n', -1.0) */
    #pragma omp parallel for collapse(2)
    for (int n = 0; n < nst; n++)
      for (int m = 0; m < nall; m++) {
        const double* cm = c + m2 * m; const double* hn = hc + m2 * n;
        double s = 0.0;
        for (size_t i = 0; i < m2; i++) s += cm[i] * hn[i];
        a[(size_t)n * nall + m] = 2.0 * s - cm[0] * hn[0];
      }
    #pragma omp parallel for
    for (int n = 0; n < nst; n++) {
      double* hn = hc + m2 * n;
      for (int m = 0; m < nall; m++) {
        const double* cm = c + m2 * m; const double am = a[(size_t)n * nall + m];
        for (size_t i = 0; i < m2; i++) hn[i] -= cm[i] * am;
      }
    }
  } else {
    /* a = c^H cp (gemm 'c','n') ; cp -= c a */
    #pragma omp parallel for collapse(2)
    for (int n = 0; n < nst; n++)
      for (int m = 0; m < nall; m++) {
        const double* cm = c + m2 * m; const double* hn = hc + m2 * n;
        double sr = 0.0, si = 0.0;
        for (int i = 0; i < ldc; i++) {
          sr += cm[2*i] * hn[2*i] + cm[2*i+1] * hn[2*i+1];
          si += cm[2*i] * hn[2*i+1] - cm[2*i+1] * hn[2*i];
        }
        a[2*((size_t)n * nall + m)] = sr; a[2*((size_t)n * nall + m)+1] = si;
      }
    #pragma omp parallel for
    for (int n = 0; n < nst; n++) {
      double* hn = hc + m2 * n;
      for (int m = 0; m < nall; m++) {
        const double* cm = c + m2 * m;
        const double ar = a[2*((size_t)n * nall + m)], ai = a[2*((size_t)n * nall + m)+1];
        for (int i = 0; i < ldc; i++) {
          hn[2*i] -= cm[2*i] * ar - cm[2*i+1] * ai;
          hn[2*i+1] -= cm[2*i] * ai + cm[2*i+1] * ar;
        }
      }
    }
  }
}

int qbo_gram(int ldc, int nst, int is_real, double* c)
{
  /* SlaterDet::gram, norm-conserving branch (SlaterDet.cc:1043-1143):
   *   real:    s = 2 c^T c (syrk 'l','t') - c(row 0)^T c(row 0) (syr) ; potrf 'l' ; c <- c L^-T  (trsm 'r','l','t','n')
   *   complex: s = c^H c (herk 'l','c') ; potrf 'l' ; c <- c L^-H (trsm 'r','l','c','n')
   * s, L: nst x nst, lower triangle, here complex for both (imaginary parts stay 0 for real bases). */
  const size_t m2 = 2 * (size_t)ldc;
  const int n = nst;
  double* s = (double*)calloc(2 * (size_t)n * n, sizeof(double));
  #pragma omp parallel for schedule(dynamic)
  for (int j = 0; j < n; j++)
    for (int i = j; i < n; i++) {                               /* s[i,j] = sum conj(c_i) c_j, i >= j */
      const double* ci = c + m2 * i; const double* cj = c + m2 * j;
      double sr = 0.0, si = 0.0;
      if (is_real) {
        for (size_t k = 0; k < m2; k++) sr += ci[k] * cj[k];
        sr = 2.0 * sr - ci[0] * cj[0];
      } else
        for (int k = 0; k < ldc; k++) {
          sr += ci[2*k] * cj[2*k] + ci[2*k+1] * cj[2*k+1];
          si += ci[2*k] * cj[2*k+1] - ci[2*k+1] * cj[2*k];
        }
      s[2*((size_t)j * n + i)] = sr; s[2*((size_t)j * n + i)+1] = si;
    }
  /* Cholesky s = L L^H, lower, column by column (LAPACK xPOTRF's definition) */
  for (int j = 0; j < n; j++) {
    double d = s[2*((size_t)j * n + j)];
    for (int k = 0; k < j; k++) { const double* l = s + 2*((size_t)k * n + j); d -= l[0]*l[0] + l[1]*l[1]; }
    if (!(d > 0.0)) { free(s); return j + 1; }
    d = sqrt(d);
    s[2*((size_t)j * n + j)] = d; s[2*((size_t)j * n + j)+1] = 0.0;
    for (int i = j + 1; i < n; i++) {
      double vr = s[2*((size_t)j * n + i)], vi = s[2*((size_t)j * n + i)+1];
      for (int k = 0; k < j; k++) {                              /* - L[i,k] conj(L[j,k]) */
        const double* li = s + 2*((size_t)k * n + i); const double* lj = s + 2*((size_t)k * n + j);
        vr -= li[0]*lj[0] + li[1]*lj[1];
        vi -= li[1]*lj[0] - li[0]*lj[1];
      }
      s[2*((size_t)j * n + i)] = vr / d; s[2*((size_t)j * n + i)+1] = vi / d;
    }
  }
  /* X L^H = C  ->  X[:,j] = ( C[:,j] - sum_{k<j} X[:,k] conj(L[j,k]) ) / L[j,j], columns in increasing order */
  for (int j = 0; j < n; j++) {
    double* xj = c + m2 * j;
    for (int k = 0; k < j; k++) {
      const double* xk = c + m2 * k;
      const double lr = s[2*((size_t)k * n + j)], li = -s[2*((size_t)k * n + j)+1];
      #pragma omp parallel for
      for (int i = 0; i < ldc; i++) {
        xj[2*i] -= xk[2*i] * lr - xk[2*i+1] * li;
        xj[2*i+1] -= xk[2*i] * li + xk[2*i+1] * lr;
      }
    }
    const double dinv = 1.0 / s[2*((size_t)j * n + j)];
    for (size_t i = 0; i < m2; i++) xj[i] *= dinv;
  }
  free(s);
  return 0;
}
