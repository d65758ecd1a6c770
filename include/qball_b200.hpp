// include/qball_b200.hpp -- C++ host-side mirror of the reference's classes on the H psi / density path, over the C ABI
// (include/qball_b200.h).  Header only.  Same names, argument meaning and error behaviour as LLNL/qball:
//
//   qb200::FourierTransform   <->  FourierTransform        (src/qball/FourierTransform.h:57-189)
//   qb200::rs_mul_add         <->  SlaterDet::rs_mul_add    (src/qball/SlaterDet.h:115, SlaterDet.cc:971-1040)
//   qb200::compute_density    <->  SlaterDet::compute_density (src/qball/SlaterDet.h:113, SlaterDet.cc:839-932)
//   qb200::NonLocalPotential  <->  NonLocalPotential::energy, norm-conserving branch (NonLocalPotential.h:93-96)
//   qb200::SubspaceLA         <->  the gemm/ger calls of PSD(A)WavefunctionStepper::update (PSDAWavefunctionStepper.cc:65-84,
//                                  264-277) and SlaterDet::gram (SlaterDet.cc:1043-1143)
//
// The reference signals errors on this path with assert / cout + MPI_Abort / exit (FourierTransform.cc:696-700,
// Messages.h:41-50); the wrappers below do the same: a non-zero status prints qb200_last_error() and aborts.
// INTEGRATION.md shows the few lines each reference class needs to forward to these.
#ifndef QBALL_B200_HPP
#define QBALL_B200_HPP
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "qball_b200.h"

namespace qb200 {

inline void check(int rc, const char* what)
{
  if (rc != QB200_OK) {
    std::fprintf(stderr, " qb200: %s failed (%d): %s\n", what, rc, qb200_last_error());
    std::abort();   // the reference: MPI_Abort(MPI_COMM_WORLD, 2)
  }
}

// What the plan needs from the reference's Basis (Basis.h:88-156); fill it from a `const Basis&` with from_basis().
struct BasisTables {
  int nrods = 0;
  std::vector<int> rod_h, rod_k, rod_lmin, rod_size;
  bool real = false;
  int idxmin1 = 0, idxmax1 = 0;
  // Works with any type exposing the reference's accessors: nrod_loc(), rod_h(i), rod_k(i), rod_lmin(i), rod_size(i),
  // real(), idxmin(1), idxmax(1).
  template <class BasisT> static BasisTables from_basis(const BasisT& b)
  {
    BasisTables t;
    t.nrods = b.nrod_loc();
    for (int i = 0; i < t.nrods; i++) {
      t.rod_h.push_back(b.rod_h(i)); t.rod_k.push_back(b.rod_k(i));
      t.rod_lmin.push_back(b.rod_lmin(i)); t.rod_size.push_back(b.rod_size(i));
    }
    t.real = b.real(); t.idxmin1 = b.idxmin(1); t.idxmax1 = b.idxmax(1);
    return t;
  }
};

class FourierTransform {
 public:
  // FourierTransform(const Basis& basis, int np0, int np1, int np2)   (FourierTransform.cc:144)
  FourierTransform(const BasisTables& b, int np0, int np1, int np2, int device = 0) : np0_(np0), np1_(np1), np2_(np2)
  {
    check(qb200_plan_create(&plan_, device, np0, np1, np2, b.nrods, b.rod_h.data(), b.rod_k.data(), b.rod_lmin.data(),
                            b.rod_size.data(), b.real ? 1 : 0, b.idxmin1, b.idxmax1), "qb200_plan_create");
  }
  ~FourierTransform() { qb200_plan_destroy(plan_); }
  FourierTransform(const FourierTransform&) = delete;
  FourierTransform& operator=(const FourierTransform&) = delete;

  // backward: Fourier synthesis, c -> f(r); forward: Fourier analysis, f is clobbered (FourierTransform.h:144-153)
  void backward(const std::complex<double>* c, std::complex<double>* f)
  { check(qb200_fft_backward(plan_, reinterpret_cast<const double*>(c), reinterpret_cast<double*>(f)), "backward"); }
  void forward(std::complex<double>* f, std::complex<double>* c)
  { check(qb200_fft_forward(plan_, reinterpret_cast<double*>(f), reinterpret_cast<double*>(c)), "forward"); }
  void backward(const std::complex<double>* c1, const std::complex<double>* c2, std::complex<double>* f)
  { check(qb200_fft_backward_pair(plan_, reinterpret_cast<const double*>(c1), reinterpret_cast<const double*>(c2), reinterpret_cast<double*>(f)), "backward(pair)"); }
  void forward(std::complex<double>* f, std::complex<double>* c1, std::complex<double>* c2)
  { check(qb200_fft_forward_pair(plan_, reinterpret_cast<double*>(f), reinterpret_cast<double*>(c1), reinterpret_cast<double*>(c2)), "forward(pair)"); }

  int np0() const { return np0_; }
  int np1() const { return np1_; }
  int np2() const { return np2_; }
  int np2_loc() const { return np2_; }                       // nprow = 1
  int np012() const { return np0_ * np1_ * np2_; }
  int np012loc() const { return np012(); }
  int index(int i, int j, int k) const { return i + np0_ * (j + np1_ * k); }   // FourierTransform.h:165
  qb200_plan* plan() const { return plan_; }
  void set_stream(void* cuda_stream) { check(qb200_plan_set_stream(plan_, cuda_stream), "set_stream"); }
  // content version of the HOST coefficient blocks passed from now on (0: always upload); see qball_b200.h
  void set_coefficient_tag(long long tag) { check(qb200_plan_set_coefficient_tag(plan_, tag), "set_coefficient_tag"); }

 private:
  qb200_plan* plan_ = nullptr;
  int np0_, np1_, np2_;
};

// sd.rs_mul_add(ft, v, sdp):  c = sd.c().cvalptr(), cp = sdp.c().valptr(), mloc = sd.c().mloc(), nstloc = sd.nstloc()
inline void rs_mul_add(FourierTransform& ft, int mloc, int nstloc, const std::complex<double>* c, const double* v,
                       std::complex<double>* cp, const double* kpg2 = nullptr)
{
  check(qb200_rs_mul_add(ft.plan(), mloc, nstloc, reinterpret_cast<const double*>(c), v, kpg2, reinterpret_cast<double*>(cp)), "rs_mul_add");
}

// sd.compute_density(ft, weight, rho):  occ_local[n] = occ_[c_.j(lj,jj)] of the n-th local state (SlaterDet.cc:912)
inline void compute_density(FourierTransform& ft, int mloc, int nstloc, const std::complex<double>* c, double weight,
                            const double* occ_local, double omega, double* rho)
{
  std::vector<double> fac(nstloc);
  const double prefac = weight / omega;                      // SlaterDet.cc:848
  for (int n = 0; n < nstloc; n++) fac[n] = prefac * occ_local[n];
  check(qb200_compute_density(ft.plan(), mloc, nstloc, reinterpret_cast<const double*>(c), fac.data(), rho), "compute_density");
}

// CurrentDensity::update_current, body of the (ispin, ikp) loop (CurrentDensity.cc:64-88): cur[idir*N + r] accumulated
inline void compute_current(FourierTransform& ft, int mloc, int nstloc, const std::complex<double>* c, double weight,
                            const double* occ_local, double omega, const double* kpgx, double* cur)
{
  std::vector<double> fac(nstloc);
  for (int n = 0; n < nstloc; n++) fac[n] = weight / omega * occ_local[n];
  check(qb200_compute_current(ft.plan(), mloc, nstloc, reinterpret_cast<const double*>(c), fac.data(), kpgx, cur), "compute_current");
}

// tail of ChargeDensity::update_density (ChargeDensity.cc:516-551): returns nelectrons_, fills rhog = vft.forward(omega*rho)
inline double density_finish(FourierTransform& vft, const double* rho, double omega, std::complex<double>* rhog)
{
  double nel = 0.0;
  check(qb200_density_finish(vft.plan(), rho, omega, reinterpret_cast<double*>(rhog), &nel), "density_finish");
  return nel;
}

// EnergyFunctional::update_vhxc + XCPotential::update on the density-basis transform (see qb200_update_vhxc); energies = exc, eps, ehart
inline void update_vhxc(FourierTransform& vft, int xc, const double* rhor, const std::complex<double>* rhog, const double* gx,
                        const double* g2i, const std::complex<double>* vion_local_g, const std::complex<double>* rhopst, double omega,
                        double* v_r, std::complex<double>* rhogt, double* energies)
{
  check(qb200_update_vhxc(vft.plan(), xc, rhor, reinterpret_cast<const double*>(rhog), gx, g2i, reinterpret_cast<const double*>(vion_local_g),
                          reinterpret_cast<const double*>(rhopst), omega, v_r, reinterpret_cast<double*>(rhogt), energies), "qb200_update_vhxc");
}

// kinetic-energy section of EnergyFunctional::energy (EnergyFunctional.cc:1155-1296) for one (spin, k-point):
// w[n] = fac * occ[c.j(lj,jj)] per local state; fills psi2sum[ngw] (may be null) and tsum[14]
inline void ekin_sums(FourierTransform& ft, int mloc, int nstloc, const std::complex<double>* c, const double* w,
                      const double* kpg2, const double* kpgx, const double* fstress, const double* dfstress, double* psi2sum,
                      double* tsum)
{
  check(qb200_ekin_sums(ft.plan(), mloc, nstloc, reinterpret_cast<const double*>(c), w, kpg2, kpgx, fstress, dfstress, psi2sum, tsum), "qb200_ekin_sums");
}

class NonLocalPotential;
// ExponentialWavefunctionStepper::exponential with a frozen Hamiltonian (ExponentialWavefunctionStepper.cc:51-149); see
// qb200_exponential.  Defined after NonLocalPotential below.
inline void exponential(FourierTransform& ft, NonLocalPotential* nlp, int mloc, int nstloc, std::complex<double>* c,
                        const double* occ_local, const double* v, const double* kpg2, double dt1, double dt2 = 0.0,
                        std::complex<double>* c2 = 0, int order = 4);

class NonLocalPotential {
 public:
  // kpgx = basis.kpgx_ptr(0) (3*ngw, component-major), omega = basis.cell().volume()
  NonLocalPotential(int ngw, bool real, double omega, const double* kpgx, int device = 0)
  { check(qb200_nl_create(&nl_, device, ngw, real ? 1 : 0, omega, kpgx), "qb200_nl_create"); }
  ~NonLocalPotential() { qb200_nl_destroy(nl_); }
  NonLocalPotential(const NonLocalPotential&) = delete;
  NonLocalPotential& operator=(const NonLocalPotential&) = delete;
  // one call per species with npr[is] > 0, in species order: na[is], npr[is], lproj[is], wt[is], twnl[is], tau[is]
  void add_species(int na, int npr, const int* lproj, const double* wt, const double* twnl, const double* tau)
  { check(qb200_nl_add_species(nl_, na, npr, lproj, wt, twnl, tau), "qb200_nl_add_species"); }
  void set_positions(int is, const double* tau) { check(qb200_nl_set_positions(nl_, is, tau), "qb200_nl_set_positions"); }
  // void NonLocalPotential::update_twnl(compute_stress = false) for Kleinman-Bylander species `is` (NonLocalPotential.cc:261-1522):
  // mproj / tabproj = m and radial-table index of every projector (iprojlm), tables = y_ / y2_ of Species::projectors_g_[l][ic] on
  // the knots gspl, gcut = the last knot of the full table.  add_species may then be given twnl = 0.
  void update_twnl(int is, const int* mproj, const int* tabproj, int ntab, int nknots, const double* gspl, double gcut,
                   const double* vnlg, const double* vnlg_spl)
  { check(qb200_nl_update_twnl(nl_, is, mproj, tabproj, ntab, nknots, gspl, gcut, vnlg, vnlg_spl), "qb200_nl_update_twnl"); }
  // semi-local species (nquad > 0): mproj[ipr] = m, rproj[ipr] = rquad[is][iquad] of projector ipr = iquad + nquad*ilm
  void update_twnl_semilocal(int is, const int* mproj, const double* rproj)
  { check(qb200_nl_update_twnl_semilocal(nl_, is, mproj, rproj), "qb200_nl_update_twnl_semilocal"); }
  // E_nl of the last energy / hpsi call (enl = 0 there: no host synchronisation); device destination: asynchronous
  void last_enl(double* enl) { check(qb200_nl_last_enl(nl_, enl), "qb200_nl_last_enl"); }
  void get_twnl(int is, double* twnl) { check(qb200_nl_get_twnl(nl_, is, twnl), "qb200_nl_get_twnl"); }
  // optional: idx = basis.idx_ptr() (3*ngw), b = { cell.b(0), cell.b(1), cell.b(2) } as 9 doubles, kpoint = basis.kpoint()
  // in crystal units -> separable phase tables; complex states at k = 0 are then contracted over the half sphere
  void set_lattice(const int* idx, const double* b, const double* kpoint)
  { check(qb200_nl_set_lattice(nl_, idx, b, kpoint), "qb200_nl_set_lattice"); }
  // double energy(SlaterDet& sd, bool compute_hpsi, SlaterDet& dsd, ...) without forces/stress
  double energy(int mloc, int nstloc, const std::complex<double>* c, const double* occ_local, bool compute_hpsi,
                std::complex<double>* cp)
  {
    double enl = 0.0;
    check(qb200_nl_energy(nl_, mloc, nstloc, reinterpret_cast<const double*>(c), occ_local, compute_hpsi ? 1 : 0,
                          reinterpret_cast<double*>(cp), &enl), "qb200_nl_energy");
    return enl;
  }
  // ---- ultrasoft beta.psi path.  The object holds, per ultrasoft species, the betag tables of
  // SlaterDet::calc_betag (twnl[lm*ngw + ig] = beta_b(|k+G|) Y_lm(k+G), lproj[lm] = l, wt unused); complex bases only.
  // void SlaterDet::calc_betapsi(): betapsi[n*M + p], p = species, atom, channel (SlaterDet.cc:2130-2263)
  void betapsi(int mloc, int nstloc, const std::complex<double>* c, std::complex<double>* betapsi)
  { check(qb200_nl_betapsi(nl_, mloc, nstloc, reinterpret_cast<const double*>(c), reinterpret_cast<double*>(betapsi)), "qb200_nl_betapsi"); }
  // cp += sum_p beta_p f_p: the gemm of calc_spsi (SlaterDet.cc:2565) / of the ultrasoft H psi with the caller's coupling in f
  void add_beta(int mloc, int nstloc, const std::complex<double>* f, std::complex<double>* cp)
  { check(qb200_nl_add_beta(nl_, mloc, nstloc, reinterpret_cast<const double*>(f), reinterpret_cast<double*>(cp)), "qb200_nl_add_beta"); }
  // void SlaterDet::calc_spsi(): spsi = c + sum beta (q <beta|psi>) / omega (SlaterDet.cc:2426-2570); qmat = per species the
  // dense symmetric npr x npr matrix of its (lm1, lm2, qaug) triples; betapsi may be null
  void spsi(int mloc, int nstloc, const std::complex<double>* c, const double* qmat, std::complex<double>* spsi,
            std::complex<double>* betapsi = 0)
  { check(qb200_nl_spsi(nl_, mloc, nstloc, reinterpret_cast<const double*>(c), qmat, reinterpret_cast<double*>(spsi),
                        reinterpret_cast<double*>(betapsi)), "qb200_nl_spsi"); }
  // ---- the rest of the ultrasoft path: tables of NonLocalPotential::update_usfns / ChargeDensity::update_usfns
  // (Species::calc_qnmg on the density basis, NonLocalPotential.cc:2719, ChargeDensity.cc:793), then the ultrasoft branch of
  // NonLocalPotential::energy (:1554-1752, no forces) and the augmentation charges of ChargeDensity::update_density (:312-465)
  void us_set_density_basis(int ngv, const double* vkpgx)
  { check(qb200_nl_us_set_density_basis(nl_, ngv, vkpgx), "qb200_nl_us_set_density_basis"); }
  void us_set_species(int is, int nq, const int* lm1, const int* lm2, const double* dzero, const std::complex<double>* qnmg)
  { check(qb200_nl_us_set_species(nl_, is, nq, lm1, lm2, dzero, reinterpret_cast<const double*>(qnmg)), "qb200_nl_us_set_species"); }
  double us_energy(int mloc, int nstloc, const std::complex<double>* c, const double* occ_local, const std::complex<double>* veff,
                   bool compute_hpsi, std::complex<double>* cp)
  {
    double enl = 0.0;
    check(qb200_nl_us_energy(nl_, mloc, nstloc, reinterpret_cast<const double*>(c), occ_local, reinterpret_cast<const double*>(veff),
                             compute_hpsi ? 1 : 0, reinterpret_cast<double*>(cp), &enl), "qb200_nl_us_energy");
    return enl;
  }
  // fac[n] = weight * occ[n] / omega as for compute_density; vft = the transform of the density basis; returns the integrated charge
  double us_augment_density(FourierTransform& vft, int mloc, int nstloc, const std::complex<double>* c, const double* fac, double* rho)
  {
    double q = 0.0;
    check(qb200_nl_us_augment_density(nl_, vft.plan(), mloc, nstloc, reinterpret_cast<const double*>(c), fac, rho, &q),
          "qb200_nl_us_augment_density");
    return q;
  }
  qb200_nl* handle() const { return nl_; }

 private:
  qb200_nl* nl_ = nullptr;
};

// the H psi block of EnergyFunctional::energy(compute_hpsi = true) in one call (EnergyFunctional.cc:1142-1153,1500,1675-1695)
inline double hpsi(FourierTransform& ft, NonLocalPotential* nlp, int mloc, int nstloc, const std::complex<double>* c,
                   const double* occ_local, const double* v, const double* kpg2, std::complex<double>* dwf)
{
  double enl = 0.0;
  check(qb200_hpsi(ft.plan(), nlp ? nlp->handle() : nullptr, mloc, nstloc, reinterpret_cast<const double*>(c), occ_local, v,
                   kpg2, reinterpret_cast<double*>(dwf), &enl), "qb200_hpsi");
  return enl;
}

inline void exponential(FourierTransform& ft, NonLocalPotential* nlp, int mloc, int nstloc, std::complex<double>* c,
                        const double* occ_local, const double* v, const double* kpg2, double dt1, double dt2,
                        std::complex<double>* c2, int order)
{
  check(qb200_exponential(ft.plan(), nlp ? nlp->handle() : nullptr, mloc, nstloc, reinterpret_cast<double*>(c), occ_local, v, kpg2,
                          order, dt1, dt2, reinterpret_cast<double*>(c2)), "qb200_exponential");
}

// Subspace dense linear algebra of the ground-state steppers, one object per SlaterDet.
class SubspaceLA {
 public:
  SubspaceLA(int ngw, bool real, int device = 0) { check(qb200_la_create(&la_, device, ngw, real ? 1 : 0), "qb200_la_create"); }
  ~SubspaceLA() { qb200_la_destroy(la_); }
  SubspaceLA(const SubspaceLA&) = delete;
  SubspaceLA& operator=(const SubspaceLA&) = delete;
  // PSDAWavefunctionStepper::update: a = c^H cp (real: 2 c^T cp - row-0 rank-1 term); cp -= c a.
  // c: mloc x n (all states), cp: mloc x nstloc (this rank's columns of H psi); a (optional): n x nstloc.
  void residual(int mloc, int n, const std::complex<double>* c, int nstloc, std::complex<double>* cp, double* a = 0)
  { check(qb200_residual(la_, mloc, n, reinterpret_cast<const double*>(c), nstloc, reinterpret_cast<double*>(cp), a), "qb200_residual"); }
  // SlaterDet::gram(): c <- c L^-H with c^H c = L L^H.  A singular overlap aborts like the reference's potrf.
  void gram(int mloc, int n, std::complex<double>* c)
  { int info = 0; check(qb200_gram(la_, mloc, n, reinterpret_cast<double*>(c), &info), "qb200_gram"); }
  // Wavefunction::diag: w (n doubles, host) = eigenvalues of c^H (H c), ascending; eigvec: c <- c z
  void diag(int mloc, int n, std::complex<double>* c, const std::complex<double>* hc, bool eigvec, double* w)
  { check(qb200_diag(la_, mloc, n, reinterpret_cast<double*>(c), reinterpret_cast<const double*>(hc), eigvec ? 1 : 0, w, 0), "qb200_diag"); }
  // the rest of PSDAWavefunctionStepper::update on device-resident blocks (see qb200_psda_update); returns theta (unclipped)
  double psda_update(qb200_comm* comm, int mloc, int nstloc, std::complex<double>* c, std::complex<double>* dc,
                     std::complex<double>* c_last, std::complex<double>* dc_last, const double* occ_local, const double* precdiag,
                     bool extrapolate)
  {
    double theta = 0.0;
    check(qb200_psda_update(la_, comm, mloc, nstloc, reinterpret_cast<double*>(c), reinterpret_cast<double*>(dc),
                            reinterpret_cast<double*>(c_last), reinterpret_cast<double*>(dc_last), occ_local, precdiag,
                            extrapolate ? 1 : 0, &theta), "qb200_psda_update");
    return theta;
  }
  // SlaterDet::gram with band-sharded states (device pointers): c_all = gathered mloc x nall block, this rank's columns
  // [first, first + nstloc) -> c_local; comm sums the overlap columns over the ranks (may be null when nstloc == nall)
  void gram_sharded(qb200_comm* comm, int mloc, int nall, const std::complex<double>* c_all, int first, int nstloc, std::complex<double>* c_local)
  {
    int info = 0;
    check(qb200_gram_sharded(la_, comm, mloc, nall, reinterpret_cast<const double*>(c_all), first, nstloc, reinterpret_cast<double*>(c_local), &info),
          "qb200_gram_sharded");
  }
  qb200_la* handle() const { return la_; }

 private:
  qb200_la* la_ = nullptr;
};

}  // namespace qb200
#endif
