/* include/qball_b200.h -- C ABI of libqball_b200.so: Qball's per-state plane-wave H psi / density hot path on B200.
 *
 * This is the drop-in boundary.  The reference (LLNL/qball, C++) has no plugin/FFI interface
 * for this path; the seam is the member functions listed beside each entry point below, which keep their signatures
 * and forward here (see INTEGRATION.md for the ~10-line shim in each reference class).  Plain pointers and sizes only.
 *
 * Conventions (identical to the reference):
 *   - complex numbers are interleaved (re,im) doubles, i.e. std::complex<double>*.
 *   - grids are x-fastest: index(i,j,k) = i + np0*(j + np1*k)          (src/qball/FourierTransform.h:165)
 *   - coefficient blocks are column-major ldc x nst with ldc = ComplexMatrix::mloc() >= ngw; rows ig >= ngw are
 *     padding, never read or written here                              (src/qball/SlaterDet.cc:2784-2787)
 *   - backward = sum_G c_G e^{+iGr}, unscaled; forward = (1/N) sum_r f e^{-iGr}  (FourierTransform.cc:1338-1342,1516-1612)
 *   - every data pointer may be a HOST pointer or a DEVICE pointer on the plan's device (detected with
 *     cudaPointerGetAttributes); host data is staged through device buffers owned by the plan, device data is used in
 *     place.  Host scalars/small tables (rod tables, occupations, fac) are always host pointers.
 *   - one handle per (spin, k-point) per rank, single-threaded callers per handle, one rank <-> one GPU
 *     (the reference's FourierTransform is stateful and not re-entrant either, FourierTransform.h:87-91).
 *   - errors: the reference aborts (assert / MPI_Abort, FourierTransform.cc:696-700); here every call returns 0 on
 *     success or a negative QB200_E* code and qb200_last_error() describes it; the C++ shim turns non-zero into
 *     message + abort.  There is NO CPU fallback: without a CUDA device every compute call fails with QB200_ENODEV.
 */
#ifndef QBALL_B200_H
#define QBALL_B200_H
#ifdef __cplusplus
extern "C" {
#endif

#define QB200_OK 0
#define QB200_EINVAL (-1)   /* bad argument                                   */
#define QB200_ENODEV (-2)   /* no usable CUDA device / kernel image           */
#define QB200_ECUDA (-3)    /* CUDA runtime error (see qb200_last_error)      */
#define QB200_ENOMEM (-4)   /* device allocation failed                       */
#define QB200_EUNSUPPORTED (-5) /* shape outside what the kernels support     */

typedef struct qb200_plan qb200_plan;

const char* qb200_last_error(void);
const char* qb200_version(void);
/* number of CUDA devices visible (0 if none / no driver); never fails */
int qb200_device_count(void);

/* ---- plan: replaces FourierTransform::FourierTransform(const Basis&, np0, np1, np2)
 *      (src/qball/FourierTransform.cc:144-526) for one rank (nprow = 1): builds the sphere<->column maps
 *      (ifftp_/ifftm_/iunpack_), ntrans0 and the FFT tables on the device.
 *      rod_* = Basis::rod_h/rod_k/rod_lmin/rod_size(irod), irod < nrods = Basis::nrod_loc()  (Basis.h:88-156)
 *      is_real = Basis::real(); idxmin1/idxmax1 = Basis::idxmin(1)/idxmax(1) (for ntrans0, FourierTransform.cc:202). */
int qb200_plan_create(qb200_plan** plan, int device, int np0, int np1, int np2, int nrods, const int* rod_h,
                      const int* rod_k, const int* rod_lmin, const int* rod_size, int is_real, int idxmin1, int idxmax1);
int qb200_plan_destroy(qb200_plan* plan);
/* work on this CUDA stream (cudaStream_t as void*); default is the legacy default stream */
int qb200_plan_set_stream(qb200_plan* plan, void* cuda_stream);
/* bytes of device scratch for the column-form intermediate (default 4 GiB when a plane fits in shared memory, 8 GiB
 * otherwise; allocated for the states actually batched); bounds the number of states per batch */
int qb200_plan_set_workspace(qb200_plan* plan, long long bytes);
/* Host coefficient blocks and residency.  When `c` is a HOST pointer, qb200_hpsi and qb200_compute_density upload it
 * in blocks of states on a copy stream while earlier blocks are being computed (and qb200_hpsi downloads finished
 * blocks of H psi on a second copy stream), so PCIe time hides behind the kernels; register the arrays with
 * cudaHostRegister (ComplexMatrix::val) -- pageable memory still works but serialises.  The device copy stays in the
 * plan.  tag != 0 declares the CONTENT VERSION of the host blocks passed from now on: a later call with the same
 * (pointer, ldc, nst) under the same tag skips the upload.  The reference evaluates rho and H psi on the same
 * wavefunction between two stepper updates (BOSampleStepper.cc: cd_.update_density(); ef_.update_vhxc(); ef_.energy()),
 * so the SlaterDet shim bumps a counter in its non-const c() accessor and passes it here.
 * tag == 0 (default): every call uploads. */
int qb200_plan_set_coefficient_tag(qb200_plan* plan, long long tag);
/* queries: 0 np0, 1 np1, 2 np2, 3 nvec, 4 ntrans0, 5 ngw, 6 is_real, 7 plane-fused path in use (1) or split path (0),
 *          8 states per batch, 9 kernels launched since creation (for bench gpu_launches),
 *          10 plane kernel in use: 0 generic, > 0 index of the compiled grid shape (plane.cu),
 *          11 second-generation z-column kernels in use, 12/13 their columns per tile (backward/forward) */
long long qb200_plan_query(const qb200_plan* plan, int what);

/* ---- FourierTransform::backward(const complex<double>* c, complex<double>* f)       FourierTransform.cc:529-539
 *      FourierTransform::forward(complex<double>* f, complex<double>* c)              FourierTransform.cc:542-552
 *      pair forms (Gamma-point two real functions per complex FFT; require is_real)   FourierTransform.cc:555-581
 *      c: ngw complex, f: np0*np1*np2 complex.  forward leaves f unspecified (the reference clobbers it too). */
int qb200_fft_backward(qb200_plan* plan, const double* c, double* f);
int qb200_fft_forward(qb200_plan* plan, double* f, double* c);
int qb200_fft_backward_pair(qb200_plan* plan, const double* c1, const double* c2, double* f);
int qb200_fft_forward_pair(qb200_plan* plan, double* f, double* c1, double* c2);

/* ---- SlaterDet::rs_mul_add(FourierTransform& ft, const double* v, SlaterDet& sdp)   SlaterDet.cc:971-1040
 *      cp[:,n] += FT[ v(r) * FT^-1[ c[:,n] ] ] for n < nst; real bases go by local pairs (n,n+1) + odd tail.
 *      kpg2 (ngw doubles, Basis::kpg2_ptr(); may be NULL) fuses the kinetic term of EnergyFunctional::energy,
 *      cp[ig,n] += 0.5*kpg2[ig]*c[ig,n]  (EnergyFunctional.cc:1675-1690; pass kpg2+fstress for confinement :1661-1668). */
int qb200_rs_mul_add(qb200_plan* plan, int ldc, int nst, const double* c, const double* v, const double* kpg2,
                     double* cp);

/* ---- SlaterDet::compute_density(FourierTransform& ft, double weight, double* rho)   SlaterDet.cc:839-932
 *      rho[i] += fac[n]*|psi_n(r_i)|^2 for every n with fac[n] > 0, fac[n] = weight*occ[n]/omega (host array).
 *      Deterministic: fixed summation order for a given (plan, nst). */
int qb200_compute_density(qb200_plan* plan, int ldc, int nst, const double* c, const double* fac, double* rho);

/* ---- tail of ChargeDensity::update_density after the sum over ranks                    ChargeDensity.cc:516-551
 *      nelectrons = sum_i rho[i] * omega / N ; rhog = FT_v[ omega * rho ]  (vft_->forward, density basis).
 *      vplan: the plan of the DENSITY basis (vbasis_: k = 0, ecut 4*ecut_wf or ecutden, ChargeDensity.cc:77-81) on the same
 *      grid; rho: N doubles; rhog: vbasis ngw complex; nelectrons may be NULL. */
int qb200_density_finish(qb200_plan* vplan, const double* rho, double omega, double* rhog, double* nelectrons);

/* ---- v(r) producers on the density basis: EnergyFunctional::update_vhxc                 EnergyFunctional.cc:353-975
 *      with XCPotential::update (XCPotential.cc:104-460) and the unpolarized LDA (Perdew-Zunger / Ceperley-Alder,
 *      LDAFunctional.cc:96-161) or PBE (PBEFunctional.cc:196-291) functional; one spin, no ESM / NLCC / enthalpy term:
 *        v_r = v_xc[rho, grad rho] + FT^-1[ vion_local_g + 4 pi (rhog/omega + rhopst) g2i ]
 *        energies[0] = exc, [1] = eps (electrons x local ionic potential), [2] = ehart
 *      vplan: the plan of the density basis (vbasis_) on the density grid; rhor: N doubles (cd_.rhor[0]); rhog: ng complex
 *      (cd_.rhog[0] = FT[omega rho]); gx = vbasis.gx_ptr(0) (3*ng, component-major; only read for PBE); g2i = vbasis.g2i_ptr();
 *      vion_local_g, rhopst: ng complex (EnergyFunctional members, rebuilt when atoms move); v_r: N doubles, OUTPUT;
 *      rhogt (may be NULL): ng complex, receives rhoelg + rhopst (needed by forces / stress).  Host or device pointers. */
#define QB200_XC_LDA 0
#define QB200_XC_PBE 1
int qb200_update_vhxc(qb200_plan* vplan, int xc, const double* rhor, const double* rhog, const double* gx, const double* g2i,
                      const double* vion_local_g, const double* rhopst, double omega, double* v_r, double* rhogt, double* energies);

/* ---- NonLocalPotential::energy, norm-conserving branch                 NonLocalPotential.cc:1909-2171, 2628-2643
 *      Projector tables are the outputs of the reference's host setup (NonLocalPotential::init/update_twnl,
 *      NonLocalPotential.cc:76-1522) and AtomSet::get_positions.  One qb200_nl per NonLocalPotential object. */
typedef struct qb200_nl qb200_nl;
int qb200_nl_create(qb200_nl** nl, int device, int ngw, int is_real, double omega, const double* kpgx /* 3*ngw */);
/* add species: na atoms, npr projectors, lproj[npr], wt[npr], twnl[npr*ngw] (twnl[is][ipr*ngw+ig]), tau[3*na] */
int qb200_nl_add_species(qb200_nl* nl, int na, int npr, const int* lproj, const double* wt, const double* twnl,
                         const double* tau);
/* NonLocalPotential::update_twnl on the device; first for a Kleinman-Bylander species (nquad == 0; NonLocalPotential.cc:261-1522, the
 * twnl part -- the stress derivatives dtwnl stay with the caller): twnl[ipr*ngw + ig] = Y_lm(k+G) * v(|k+G|), real spherical
 * harmonics in the reference's order and normalisation, v from the species' radial cubic splines (Species::dvnlg,
 * Species.cc:1492-1505; splintd, spline.cc:126-156).  After a cell change the caller refreshes kpgx (a new object) or the
 * tables with this call instead of shipping npr*ngw doubles.
 *   mproj[ipr]   = m of projector ipr (iprojlm[is][l][m][ic] == ipr, NonLocalPotential.cc:219-233), 0 <= m <= 2 l
 *   tabproj[ipr] = index of its radial table, one per (l, channel): vnlg / vnlg_spl [ntab][nknots] = the y_ / y2_ arrays of
 *                  Species::projectors_g_[l][ic] on the knots gspl[nknots] (the first nknots of Species::gspl_; enough knots to
 *                  cover max |k+G| suffice), gcut = the LAST knot of the full table (beyond it v = 0, Species.cc:1495)
 * qb200_nl_add_species accepts twnl == NULL for a species whose table is filled this way.  All pointers here are host pointers.
 * qb200_nl_get_twnl copies the table of species `is` (npr*ngw doubles) back, host or device destination. */
int qb200_nl_update_twnl(qb200_nl* nl, int is, const int* mproj, const int* tabproj, int ntab, int nknots, const double* gspl,
                         double gcut, const double* vnlg, const double* vnlg_spl);
/* the same for a semi-local species (nquad > 0; NonLocalPotential.cc:366-419, 500-600, 800-960, 1230-1345): projector
 * ipr = iquad + nquad*ilm has twnl[ipr*ngw + ig] = Y_lm(k+G) * 4 pi j_l(|k+G| r) r at its quadrature radius
 * rproj[ipr] = rquad[is][iquad] (NonLocalPotential.cc:150-200); mproj[ipr] = m.  Host pointers. */
int qb200_nl_update_twnl_semilocal(qb200_nl* nl, int is, const int* mproj, const double* rproj);
int qb200_nl_get_twnl(qb200_nl* nl, int is, double* twnl);
/* optional, before the first energy call: the integer description of the plane waves.
 *   idx    = Basis::idx_ptr(): 3*ngw ints, (h,k,l) of plane wave ig at idx[3*ig + 0..2]      (Basis.cc:672-674)
 *   b      = reciprocal lattice vectors UnitCell::b(0), b(1), b(2) as 9 doubles (b[3*i + xyz]) (Basis.cc:711-713)
 *   kpoint = Basis::kpoint(), in units of b0,b1,b2                                             (Basis.h:51)
 * With k+G = kpoint + h b0 + k b1 + l b2 the structure-factor phase exp(-i (k+G).tau) (comp_eigr / comp_anl,
 * NonLocalPotential.cc:1959-2036) factorises per direction; the kernels then read three small per-atom tables instead
 * of evaluating one FP64 sincos per (atom, G) and tile.  Same results to rounding (~1e-13 absolute in the phase).
 * With this description, a COMPLEX basis at kpoint = 0 (force_complex_wf, the TDDFT configuration) whose tables satisfy
 * twnl(-G) = (-1)^l twnl(G) (every table NonLocalPotential::update_twnl produces; verified on the device) is contracted
 * over the half sphere: psi = psi_R + i psi_I as two real functions, a third fewer flops than the general complex path
 * (qb200_nl_query(nl, 14) == 3 after an energy call; QB200_NL_GAMMA=0 disables it). */
int qb200_nl_set_lattice(qb200_nl* nl, const int* idx, const double* b, const double* kpoint);
/* atoms moved: new positions for species is (AtomSet::get_positions order) */
int qb200_nl_set_positions(qb200_nl* nl, int is, const double* tau);
int qb200_nl_set_stream(qb200_nl* nl, void* cuda_stream);
/* bytes of device memory the materialised anl block may take (default 8 GiB).  anl for all projectors and a chunk of
 * plane waves is written once per energy call (the reference's comp_anl) in the layout both GEMMs read; when the whole
 * sphere does not fit (Au992: 181 GB) the call sweeps over chunks of plane waves instead of blocks of atoms. */
int qb200_nl_set_workspace(qb200_nl* nl, long long bytes);
int qb200_nl_destroy(qb200_nl* nl);
/* enl = sum_{n,I,p} occ[n]*wt_p/omega*|F_{Ip,n}|^2 ; if compute_hpsi: cp += anl * (wt/omega * F).
 * occ: host array of the nst LOCAL states' occupations (occ[c.j(lj,jj)] in the reference, NonLocalPotential.cc:2115).
 * The row-sum of enl over G-row ranks (NonLocalPotential.cc:2629) is the identity with nprow = 1. */
int qb200_nl_energy(qb200_nl* nl, int ldc, int nst, const double* c, const double* occ, int compute_hpsi, double* cp,
                    double* enl);
/* E_nl of the last qb200_nl_energy / qb200_hpsi call.  A caller that passes enl = NULL to those calls gets no host
 * synchronisation from them; the energy stays in a device scalar that this call copies out: to a DEVICE address without
 * synchronising (on the object's stream; e.g. behind the density, so that one qb200_allreduce_rho carries rho and E_nl --
 * the fused small all-reduce of NonLocalPotential.cc:2629 / ChargeDensity.cc:309), or to a HOST address (synchronous). */
int qb200_nl_last_enl(qb200_nl* nl, double* enl);
/* ---- Ultrasoft beta.psi path.  The object is created as above with, per species, the reference's
 * betag tables -- twnl[lm*ngw + ig] = beta_b(|k+G|) * Y_lm(k+G), lproj[lm] = l of channel lm, wt unused -- as
 * SlaterDet::calc_betag fills them before the (-i)^l factor (src/qball/SlaterDet.cc:2006-2127; an input like twnl).  Complex
 * bases only: ultrasoft potentials force complex states (SlaterDet.cc:57-58); QB200_EUNSUPPORTED otherwise.
 * Projector order p: species in the order added, then atom, then channel (ia*npr + lm), M = total number.
 *   qb200_nl_betapsi : betapsi[n*M + p] = sum_G conj(anl_p(G)) c_n(G), anl = betag * (-i)^l * exp(-i (k+G).tau)
 *                      replaces SlaterDet::calc_betapsi (SlaterDet.cc:2130-2263)
 *   qb200_nl_add_beta: cp_n(G) += sum_p anl_p(G) f[n*M + p]            the gemm of SlaterDet::calc_spsi (:2565) and of the
 *                      ultrasoft H psi (NonLocalPotential.cc:1554-1906) with the caller's coupling already applied to f
 *   qb200_nl_spsi    : spsi = c + sum anl_p (q betapsi)_p / omega; qmat = per species, in order, the dense symmetric npr x npr
 *                      matrix built from the (lm1, lm2, qaug) triples of SlaterDet::calc_spsi (:2536-2549); betapsi (optional,
 *                      may be NULL) receives <beta|psi>.  replaces SlaterDet::calc_spsi (SlaterDet.cc:2426-2570)
 * c, cp, spsi: ldc x nst complex blocks, host or device; betapsi, f: nst x M complex, host or device; qmat: host or device. */
int qb200_nl_betapsi(qb200_nl* nl, int ldc, int nst, const double* c, double* betapsi);
int qb200_nl_add_beta(qb200_nl* nl, int ldc, int nst, const double* f, double* cp);
int qb200_nl_spsi(qb200_nl* nl, int ldc, int nst, const double* c, const double* qmat, double* spsi, double* betapsi);
/* ---- Ultrasoft potentials, the rest: the ultrasoft branch of NonLocalPotential::energy
 * (src/qball/NonLocalPotential.cc:1554-1752; forces :1754-1900 stay with the caller) and the augmentation charges of
 * ChargeDensity::update_density (src/qball/ChargeDensity.cc:312-465).  Same object as above (betag tables per species).
 * Further inputs, all produced by the reference's host setup like twnl / betag:
 *   qb200_nl_us_set_density_basis: the density basis cdbasis_ / vbasis_ (complex at k = 0 for ultrasoft runs,
 *       ChargeDensity.cc:75): ngv = localsize(), vkpgx = kpgx_ptr(0) (3*ngv, component-major).  Before set_species.
 *   qb200_nl_us_set_species: species `is` (order of qb200_nl_add_species): nq = Species::nqtot() pairs of channels
 *       lm1[q] = qnm_lm1(q), lm2[q] = qnm_lm2(q), dzero[q] = Species::dzero(q), qnmg[q*ngv + ig] = Q_q(G) complex as
 *       Species::calc_qnmg(cdbasis_, .) returns it (NonLocalPotential.cc:2719, ChargeDensity.cc:793; no structure factor).
 *   qb200_nl_us_energy: D^I_q = dzero_q + sum_G Re(conj(exp(-i G.tau_I) Q_q(G)) veff(G))          (:1607-1636)
 *       *enl = sum_n occ_n/omega sum_{I,q} mult_q dzero_q Re(conj(bp_n[I,lm1]) bp_n[I,lm2])          (:1639-1665)
 *       compute_hpsi: cp_n += sum_{I,lm} beta^I_lm(G) (1/omega) sum_lm' D^I[lm,lm'] bp_n[I,lm']      (:1667-1750)
 *       veff: ngv complex (veff_g of EnergyFunctional.cc:924-927), only read when compute_hpsi; occ: nst doubles.
 *   qb200_nl_us_augment_density: rho += Re FT^-1[ sum_{I,q} Q_q(G) summat[I,q] exp(-i G.tau_I)/omega ],
 *       summat[I,q] = sum_n fac_n mult_q conj(bp_n[I,lm1]) bp_n[I,lm2], fac_n = weight*occ_n/omega as for
 *       qb200_compute_density; vplan = the plan of the density basis on the density grid; *uscharge (may be NULL) receives
 *       the integral the reference prints (:441-446).  The term is a sum over states, so a band-sharded caller adds each
 *       rank's part to its partial rho before qb200_allreduce_rho (the reference sums summat over the ranks instead, :373-375).
 * Pointers host or device.  Complex bases only (QB200_EUNSUPPORTED otherwise). */
int qb200_nl_us_set_density_basis(qb200_nl* nl, int ngv, const double* vkpgx);
int qb200_nl_us_set_species(qb200_nl* nl, int is, int nq, const int* lm1, const int* lm2, const double* dzero, const double* qnmg);
int qb200_nl_us_energy(qb200_nl* nl, int ldc, int nst, const double* c, const double* occ, const double* veff, int compute_hpsi,
                       double* cp, double* enl);
int qb200_nl_us_augment_density(qb200_nl* nl, qb200_plan* vplan, int ldc, int nst, const double* c, const double* fac, double* rho,
                                double* uscharge);
long long qb200_nl_query(const qb200_nl* nl, int what); /* 9: kernels launched, 11: plane-wave chunks of the last call,
                                                           12: bytes of the anl block, 13: projectors in total,
                                                           14: form of the last call: 0 real basis, 1 complex 4-product,
                                                               2 complex 3-product, 3 Gamma-point half sphere */

/* ---- the whole H psi column block in the reference's order (EnergyFunctional.cc:1142-1153, 1500, 1675-1695):
 *      hpsi = 0 ; hpsi += V_nl psi ; hpsi += 0.5|k+G|^2 psi ; hpsi += FT[v FT^-1 psi].  nl may be NULL (no projectors).
 *      hpsi is OUTPUT only (ldc x nst); enl may be NULL. */
int qb200_hpsi(qb200_plan* plan, qb200_nl* nl, int ldc, int nst, const double* c, const double* occ, const double* v,
               const double* kpg2, double* hpsi, double* enl);

/* ---- kinetic-energy sums of EnergyFunctional::energy                               EnergyFunctional.cc:1155-1296
 *      psi2sum[ig] = sum_n w[n] |c[ig,n]|^2 with w[n] = fac * occ[n] (host array of the nst LOCAL states; fac = 1 for a real
 *      basis, 0.5 otherwise, :1184), then the 14 partial sums of :1225-1276 over this rank's plane waves:
 *        tsum[0] = sum psi2sum*kpg2 (ekin), tsum[1..6] = sum 2 psi2sum*(xx,yy,zz,xy,yz,xz) (sigma_ekin; only if kpgx != NULL),
 *        tsum[7] = sum psi2sum*fstress (econf; if fstress != NULL), tsum[8..13] = sum psi2sum*dfstress*(xx,..) (if both).
 *      kpg2 = Basis::kpg2_ptr() (ngw), kpgx = Basis::kpgx_ptr(0) (3*ngw, component-major) or NULL,
 *      fstress/dfstress = ConfinementPotential::fstress()/dfstress() or NULL.  psi2sum (ngw doubles) may be NULL.
 *      tsum: 14 doubles, host or device.  The k-point weight, 1/weightsum and the dsum over ranks (:1281-1294) stay with the
 *      caller.  A host block whose coefficient tag is unchanged is not uploaded again.  Deterministic. */
int qb200_ekin_sums(qb200_plan* plan, int ldc, int nst, const double* c, const double* w, const double* kpg2,
                    const double* kpgx, const double* fstress, const double* dfstress, double* psi2sum, double* tsum);

/* ---- TDDFT propagator glue: ExponentialWavefunctionStepper::exponential(num_exp, dt1, dt2)
 *      (ExponentialWavefunctionStepper.cc:51-149) with the Hamiltonian frozen at v (the caller updates v between the
 *      exponentials of an ETRS/AETRS step): c <- sum_{N=0..order} (-i dt1 H)^N / N! c  (order 4 in the reference, :45);
 *      if c2 != NULL (num_exp == 2): c2 <- the same series with dt2, from the same H^N c.  H = qb200_hpsi.
 *      Complex bases only (the reference requires force_complex_wf for wf_dyn ETRS, vars/WfDyn.h:82-92).
 *      The H^N c chain stays on the device; work space: two blocks of ldc x nst complex owned by the plan. */
int qb200_exponential(qb200_plan* plan, qb200_nl* nl, int ldc, int nst, double* c, const double* occ, const double* v,
                      const double* kpg2, int order, double dt1, double dt2, double* c2);

/* ---- CurrentDensity::update_current (CurrentDensity.cc:52-86) for one spin / k-point: the pair density of
 *      SlaterDet::compute_density(FourierTransform&, double weight, complex<double>* rho, const SlaterDet& sd2)
 *      (SlaterDet.cc:935-968) with sd2 = i*kpgx[idir]*c (:72-76), reduced to what the caller keeps:
 *        cur[idir*N + r] += -Im sum_n fac[n] conj(psi_n(r)) FT^-1[i kpgx_idir c_n](r)      for fac[n] > 0, idir = 0,1,2
 *      kpgx = Basis::kpgx_ptr(0) (3*ngw doubles, component-major), fac[n] = weight*occ[n]/omega (host array, as in
 *      qb200_compute_density), cur: 3*N doubles, ACCUMULATED (the caller clears it and sums over k-points and ranks,
 *      CurrentDensity.cc:60-62, 90).  Real bases add nothing (real wavefunctions carry no current). */
int qb200_compute_current(qb200_plan* plan, int ldc, int nst, const double* c, const double* fac, const double* kpgx,
                          double* cur);

/* ---- subspace dense linear algebra between two H psi evaluations, on the same FP64
 *      tensor-core GEMM kernels as the projector contractions.  One qb200_la per (spin, k-point) wavefunction block.
 *      Blocks are ComplexMatrix::val as everywhere else (column-major ldc x n, rows >= ngw padding); real bases are read
 *      through the reference's DoubleMatrix proxy view (2*ldc real rows, PSDAWavefunctionStepper.cc:67-69). */
typedef struct qb200_la qb200_la;
int qb200_la_create(qb200_la** la, int device, int ngw, int is_real);
int qb200_la_set_stream(qb200_la* la, void* cuda_stream);
/* bytes the packed copy of the wavefunction block may take (default 8 GiB); larger blocks are swept in plane-wave chunks */
int qb200_la_set_workspace(qb200_la* la, long long bytes);
int qb200_la_destroy(qb200_la* la);
long long qb200_la_query(const qb200_la* la, int what); /* 9: kernels launched, 11: plane-wave chunks of the last call,
                                                           13: the last qb200_diag replayed its Jacobi sweeps from a CUDA graph */
/* descent direction of the SD / PSD / PSDA wavefunction steppers:
 *   complex:  a.gemm('c','n',1.0,c,cp,0.0); cp.gemm('n','n',-1.0,c,a,1.0)       PSDAWavefunctionStepper.cc:264-277
 *   real:     a.gemm('t','n',2.0,c,cp,0.0); a.ger(-1.0,c,0,cp,0); cp.gemm('n','n',-1.0,c,a,1.0)          :65-84
 *   (PSDWavefunctionStepper.cc:62-90 is the same sequence)
 * c:  ldc x nall, ALL states of the Slater determinant (with band sharding: the gathered block, qball_b200/parallel.py);
 * hc: ldc x nst, this rank's columns of H psi, replaced by H psi - c a;
 * a:  optional output, nall x nst column-major (complex, or real for real bases) = this rank's columns of a. */
int qb200_residual(qb200_la* la, int ldc, int nall, const double* c, int nst, double* hc, double* a);
/* SlaterDet::gram(), norm-conserving branch                                               SlaterDet.cc:1043-1143
 *   complex: s.herk('l','c',1.0,c,0.0); s.potrf('l'); c.trsm('r','l','c','n',1.0,s)
 *   real:    s.syrk('l','t',2.0,c,0.0); s.syr('l',-1.0,c,0,'r'); s.potrf('l'); c.trsm('r','l','t','n',1.0,s)
 * c: ldc x nst (all states on this rank), orthonormalised in place.  info (may be NULL): LAPACK potrf's info; a
 * non-positive-definite overlap returns QB200_EINVAL with *info = order of the failing minor and leaves c unchanged. */
int qb200_gram(qb200_la* la, int ldc, int nst, double* c, int* info);
/* SlaterDet::gram with the states sharded over ranks (the reference's distributed herk / potrf / trsm, SlaterDet.cc:1043-1143,
 * restated for band sharding with nprow = 1).  c_all: the gathered ldc x nall block (as for qb200_residual); this rank owns the
 * columns [first, first + nst); c_local (ldc x nst; may be those columns of c_all) receives the orthonormalised states.
 *   qb200_gram_overlap: S (nall x nall complex, column-major) = 0 except this rank's columns S[:, first..] = c_all^H c_local
 *   (sum S over the ranks: every entry has one non-zero contributor)
 *   qb200_gram_apply  : Cholesky S = L L^H (replicated on every rank), c_local <- c_all (L^-H)[:, first..]
 *   qb200_gram_sharded: the three steps with the sum done by qb200_allreduce_rho on `comm` (NULL allowed when nst == nall)
 * Device pointers only.  info as for qb200_gram. */
int qb200_gram_overlap(qb200_la* la, int ldc, int nall, const double* c_all, int first, int nst, double* S);
int qb200_gram_apply(qb200_la* la, int ldc, int nall, const double* c_all, const double* S, int first, int nst, double* c_local, int* info);
int qb200_gram_sharded(qb200_la* la, struct qb200_comm* comm, int ldc, int nall, const double* c_all, int first, int nst, double* c_local, int* info);

/* ---- Wavefunction::diag(dwf, eigvec), one (spin, k-point), norm-conserving                     Wavefunction.cc:1510-1715
 *      h = c^H (H c) (real bases: 2 c^T (H c) minus the rank-1 term of real row 0, :1538-1539); w = eigenvalues of h from its
 *      lower triangle, ascending (syevd / heevd 'l', :1604, :1682; sd->set_eig(w)); eigvec != 0: c <- c z with z the
 *      eigenvectors (:1606-1609, :1684-1688).  The reference calls (Sca)LAPACK; here: parallel cyclic Jacobi on the device
 *      + the FP64 tensor-core GEMM for c z.  c: ldc x nst (all states on this rank), overwritten only if eigvec; hc = H c,
 *      ldc x nst (read only); w: nst doubles, HOST; sweeps (may be NULL): Jacobi sweeps used. */
int qb200_diag(qb200_la* la, int ldc, int nst, double* c, const double* hc, int eigvec, double* w, int* sweeps);

/* ---- optional per-kernel timing (CUDA events on the launching stream, recorded around every launch while enabled).
 *      categories: 0 k_zcol_bwd, 1 xy stage (k_plane, or k_xrows+k_ycols), 2 k_zcol_fwd, 3 k_fnl, 4 k_fnl_finish+sum,
 *      5 k_back, 6 k_rho_reduce, 7 k_anl_gen.  qb200_profile_read synchronises, ADDS elapsed milliseconds and launch counts of the
 *      recorded launches into ms[ncat]/count[ncat] and clears the record. */
#define QB200_NCAT 8
int qb200_profile_enable(int on);
int qb200_profile_read(double* ms, long long* count, int ncat);

/* ---- the path's collectives over the GPUs of one box (band parallelism, nprow = 1, one rank per GPU), NCCL inside:
 *      qb200_allreduce_rho      replaces  wfcontext->dsum('r', np012loc, 1, &rhor[ispin][0], np012loc)      ChargeDensity.cc:309
 *      qb200_allreduce_scalars  replaces  ctxt_.dsum('r',1,1,&enl,1)   NonLocalPotential.cc:2629,
 *                                         wfcontext()->dsum(14,1,&sum[0],14)  EnergyFunctional.cc:1294 and the nelectrons
 *                                         dsum of ChargeDensity.cc:528 (pack them into one call: {E_nl, tsum[14], integral of rho})
 *      Setup: rank 0 calls qb200_comm_get_unique_id, the caller broadcasts the 128 bytes out of band (MPI_Bcast over the
 *      reference's own communicator), every rank calls qb200_comm_init(device = its GPU, id, rank, nranks).
 *      rho: n doubles, DEVICE pointer (summed in place, enqueued on `stream` like a kernel launch) or HOST pointer (staged,
 *      synchronous).  vals: a short HOST array.  NCCL is loaded at run time (libnccl.so.2); without it these calls return
 *      QB200_EUNSUPPORTED and everything else keeps working. */
#define QB200_UNIQUE_ID_BYTES 128
typedef struct qb200_comm qb200_comm;
int qb200_comm_get_unique_id(void* id /* QB200_UNIQUE_ID_BYTES */);
int qb200_comm_init(qb200_comm** comm, int device, const void* id, int rank, int nranks);
int qb200_comm_destroy(qb200_comm* comm);
long long qb200_comm_query(const qb200_comm* comm, int what);   /* 0 rank, 1 nranks, 2 NCCL version code */
int qb200_allreduce_rho(qb200_comm* comm, double* rho, long long n, void* cuda_stream);
int qb200_allreduce_scalars(qb200_comm* comm, double* vals, int n);

/* ---- the rest of PSDAWavefunctionStepper::update after the descent direction (qb200_residual), so that the wavefunction
 *      block stays on the device between two H psi evaluations          PSDAWavefunctionStepper.cc:93-225 (real), :281-395
 *      with Preconditioner::apply(sd, ispin, ikp, -1.0)                   Preconditioner.cc:118-139
 *        dc[ig,n] *= -precdiag[ig] (ig < ngw; precdiag = Preconditioner::diag(ispin, ikp), Preconditioner.cc:47-90)
 *        extrapolate != 0 (every call but the first after a reset, extrapolate_[ispin][ikp]):
 *          a = sum occ_n f (f - f_last), b = sum occ_n (f - f_last)^2 over the local columns (real basis: G and -G counted,
 *          G = 0 once), summed over the ranks of `comm` (may be NULL: one rank); theta = -a/b; theta < -1 -> 0; min(2, theta)
 *          c <- c + theta (c - c_last) + f + theta (f - f_last)
 *        extrapolate == 0: c <- c + f.        In both cases c_last <- old c, dc_last <- f.
 *      c, dc, c_last, dc_last: ldc x nst DEVICE blocks (wf_, dwf, wf_last_, dwf_last_); occ: host, the nst local occupations;
 *      precdiag: ngw doubles, host or device.  *theta (may be NULL) receives -a/b before clipping (the value the reference
 *      prints).  The orthogonalisation that follows in the reference (SlaterDet::gram) is qb200_gram. */
int qb200_psda_update(qb200_la* la, qb200_comm* comm, int ldc, int nst, double* c, double* dc, double* c_last, double* dc_last,
                      const double* occ, const double* precdiag, int extrapolate, double* theta);

/* ---- device self-measurement for the FP64 roofline denominator (bench.py): out[0] = FP64 tensor (DMMA, mma.sync.m8n8k4.f64)
 *      TFLOP/s, out[1] = plain DFMA TFLOP/s, issue-rate loops on every SM of `device` (~50 ms).  The reference has no
 *      counterpart; MEASURED_PEAKS.json carries no FP64 figure. */
int qb200_measure_fp64_peak(int device, double* out);

#ifdef __cplusplus
}
#endif
#endif
