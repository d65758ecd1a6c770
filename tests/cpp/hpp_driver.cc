// tests/cpp/hpp_driver.cc -- exercises include/qball_b200.hpp the way the reference's classes would call it
// (host std::complex arrays in, host arrays out).  Built and run by tests/test_gpu_cpp_mirror.py.
#include <qball_b200.hpp>
#include <cstdio>
#include <vector>

struct FakeBasis {  // the accessors of the reference's Basis that BasisTables::from_basis uses (Basis.h:88-156)
  int nrods_, real_, imin1, imax1;
  std::vector<int> h, k, lmin, size;
  int nrod_loc() const { return nrods_; }
  int rod_h(int i) const { return h[i]; }
  int rod_k(int i) const { return k[i]; }
  int rod_lmin(int i) const { return lmin[i]; }
  int rod_size(int i) const { return size[i]; }
  bool real() const { return real_ != 0; }
  int idxmin(int) const { return imin1; }
  int idxmax(int) const { return imax1; }
};

template <class T> static void rd(FILE* f, T* p, size_t n) { if (fread(p, sizeof(T), n, f) != n) { fprintf(stderr, "short read\n"); exit(2); } }

int main(int argc, char** argv)
{
  if (argc < 3) return 1;
  FILE* f = fopen(argv[1], "rb");
  if (!f) return 1;
  int hd[10];
  rd(f, hd, 10);
  const int np0 = hd[0], np1 = hd[1], np2 = hd[2], nrods = hd[3], ngw = hd[7], ldc = hd[8], nst = hd[9];
  FakeBasis b; b.nrods_ = nrods; b.real_ = hd[4]; b.imin1 = hd[5]; b.imax1 = hd[6];
  b.h.resize(nrods); b.k.resize(nrods); b.lmin.resize(nrods); b.size.resize(nrods);
  rd(f, b.h.data(), nrods); rd(f, b.k.data(), nrods); rd(f, b.lmin.data(), nrods); rd(f, b.size.data(), nrods);
  const size_t N = (size_t)np0 * np1 * np2;
  std::vector<std::complex<double> > c((size_t)ldc * nst), cp((size_t)ldc * nst), fr(N);
  std::vector<double> v(N), kpg2(ngw), occ(nst), rho(N, 0.0);
  double omega;
  rd(f, c.data(), c.size()); rd(f, v.data(), N); rd(f, kpg2.data(), ngw); rd(f, occ.data(), nst); rd(f, &omega, 1);
  fclose(f);
  qb200::FourierTransform ft(qb200::BasisTables::from_basis(b), np0, np1, np2);
  ft.backward(&c[0], &fr[0]);
  qb200::rs_mul_add(ft, ldc, nst, c.data(), v.data(), cp.data(), kpg2.data());
  qb200::compute_density(ft, ldc, nst, c.data(), 1.0, occ.data(), omega, rho.data());
  // the stepper's descent direction on (c, cp) and SlaterDet::gram on c, as PSDAWavefunctionStepper / SlaterDet would call
  std::vector<std::complex<double> > res(cp), cg(c);
  qb200::SubspaceLA la(ngw, b.real());
  la.residual(ldc, nst, c.data(), nst, res.data());
  la.gram(ldc, nst, cg.data());
  FILE* o = fopen(argv[2], "wb");
  fwrite(fr.data(), 16, N, o); fwrite(cp.data(), 16, cp.size(), o); fwrite(rho.data(), 8, N, o);
  fwrite(res.data(), 16, res.size(), o); fwrite(cg.data(), 16, cg.size(), o);
  fclose(o);
  return 0;
}
