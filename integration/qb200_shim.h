// integration/qb200_shim.h -- what the reference's seam functions call when Qball is built with -DUSE_QB200
// (INTEGRATION.md).  The reference's class declarations are NOT changed: the device objects that belong to a
// FourierTransform / NonLocalPotential live in side tables keyed by the object's address, so only the .cc files of the
// seam are recompiled (integration/apply_shim.py inserts the forwarding lines; oracle/Makefile target `ref_qb200`).
//
// Run-time switch: QB200_SHIM=0 leaves every call on the reference's CPU path (used by the parity tests to produce the
// unmodified numbers from the same binary); otherwise a CUDA device is REQUIRED -- there is no silent fallback.
#ifndef QB200_SHIM_H
#define QB200_SHIM_H
#include <complex>
#include <vector>

class FourierTransform;
class NonLocalPotential;
class Basis;
class AtomSet;
class SlaterDet;
class ComplexMatrix;

namespace qb200 { class FourierTransform; }

namespace qb200_shim {

bool enabled();
// FourierTransform::FourierTransform / ~FourierTransform  (FourierTransform.cc:144, :114)
void ft_attach(const FourierTransform* key, const Basis& basis, int np0, int np1, int np2);
void ft_detach(const FourierTransform* key);
qb200::FourierTransform& ft_gpu(const FourierTransform* key);
// SlaterDet::rs_mul_add / compute_density (SlaterDet.cc:971, :839); occ_loc[n] = occ_[c_.j(lj,jj)]
void rs_mul_add(const FourierTransform& ft, const ComplexMatrix& c, int nstloc, const double* v, ComplexMatrix& cp);
void compute_density(const FourierTransform& ft, const ComplexMatrix& c, int nstloc, const std::vector<double>& occ, double weight,
                     double omega, double* rho);
// NonLocalPotential::energy, norm-conserving branch without forces / stress (NonLocalPotential.cc:1909-2171)
double nl_energy(const NonLocalPotential* key, const Basis& basis, AtomSet& atoms, int nsp, const std::vector<int>& na,
                 const std::vector<int>& npr, const std::vector<std::vector<int> >& lproj,
                 const std::vector<std::vector<double> >& wt, const std::vector<std::vector<double> >& twnl, SlaterDet& sd,
                 bool compute_hpsi, SlaterDet& dsd);
void nl_invalidate(const NonLocalPotential* key);      // update_twnl rebuilt the tables / the object dies
// EnergyFunctional::energy, psi2sum loop (EnergyFunctional.cc:1209-1223): returns true when psi2sum was filled on the device
bool psi2sum(const FourierTransform* ft, const ComplexMatrix& c, const double* occ_global, double fac, const double* kpg2,
             std::vector<double>& psi2sum);
// EnergyFunctional::update_vhxc (EnergyFunctional.cc:353-975) for one spin, LDA (xc 0) / PBE (xc 1); energies = exc, eps, ehart
void update_vhxc(const FourierTransform* vft, int xc, const double* rhor, const std::complex<double>* rhog, const double* gx,
                 const double* g2i, const std::complex<double>* vion_local_g, const std::complex<double>* rhopst, double omega,
                 double* v_r, std::complex<double>* rhogt, double* energies);
// counters for the tests: calls forwarded to the device since start
long long forwarded_calls();

}  // namespace qb200_shim
#endif
