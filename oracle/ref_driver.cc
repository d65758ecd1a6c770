// oracle/ref_driver.cc -- TEST INFRASTRUCTURE ONLY.
//
// Drives the UNMODIFIED reference classes (linked from oracle/_ref/libqballref.a, compiled from the sources where
// they lie under /root/reference) on explicit input arrays and dumps raw arrays, so that the C restatement
// (oracle/qb_oracle.c) and the CUDA path can be pinned against the reference itself at array level.
//
// The reference calls exercised (all on one rank, nprow = npcol = 1):
//   Basis::resize                         src/qball/Basis.cc:302-700
//   FourierTransform::backward/forward    src/qball/FourierTransform.cc:529-581 (+ pair forms)
//   SlaterDet::rs_mul_add                 src/qball/SlaterDet.cc:971-1040
//   SlaterDet::compute_density            src/qball/SlaterDet.cc:839-932
//   NonLocalPotential::energy (NC branch) src/qball/NonLocalPotential.cc:1909-2171
//   DoubleMatrix/ComplexMatrix gemm, ger  as called by PSDAWavefunctionStepper::update (PSDAWavefunctionStepper.cc:65-84, 264-277)
//   SlaterDet::gram                       src/qball/SlaterDet.cc:1043-1143
// The kinetic term of EnergyFunctional::energy (EnergyFunctional.cc:1675-1690) lives inside a function that needs a
// whole Sample; its three-line loop is restated here in the same order (clear -> nonlocal -> kinetic -> local).
//
//   NonLocalPotential::energy (ultrasoft branch) src/qball/NonLocalPotential.cc:1554-1906   } mode `usx`: through a
//   ChargeDensity::update_density (augmentation) src/qball/ChargeDensity.cc:312-465         } Sample, as the application does
// usage:  ref_driver basis <case.txt>     dump basis / grid tables to <out>.*
//         ref_driver run   <case.txt>     read <out>.in_c.f64, <out>.in_v.f64, <out>.in_occ.f64 ; dump results
//
// case file (one directive per line):
//   cell a0x a0y a0z a1x a1y a1z a2x a2y a2z     (bohr)
//   ecut E                                        (hartree)
//   kpoint kx ky kz                               (crystal units, as the reference's `kpoint` command)
//   force_complex 0|1
//   grid np0 np1 np2                              (0 0 0 => ChargeDensity::initialize rule, ChargeDensity.cc:77-99)
//   nst N
//   species <name> <file.xml>
//   atom <name> <species> x y z                   (bohr)
//   out <prefix>

#include <fstream>
#include <sstream>
#include <iostream>
#include <iomanip>
#include <vector>
#include <valarray>
#include <complex>
#include <string>
#include <map>
#include <list>
#include <cstdio>
#include <cstring>
#include <omp.h>
// the driver needs tables that are private members of the reference's classes (NonLocalPotential: twnl, wt, lproj, iprojlm;
// Species / Spline: the radial spline tables behind Species::dvnlg).  Every standard header is included above, so only the
// reference's own classes are affected (access specifiers do not change the layout).
#define private public
#include <qball/Basis.h>
#include <qball/FourierTransform.h>
#include <qball/SlaterDet.h>
#include <qball/AtomSet.h>
#include <qball/Atom.h>
#include <qball/Species.h>
#include <qball/SpeciesReader.h>
#include <qball/Context.h>
#include <qball/UnitCell.h>
#include <qball/Timer.h>
#include <qball/VectorPotential.h>
#include <qball/Sample.h>
#include <qball/Wavefunction.h>
#include <qball/ChargeDensity.h>
#include <math/matrix.h>
#include <functionals/LDAFunctional.h>
#include <functionals/PBEFunctional.h>
#include <qball/NonLocalPotential.h>
#undef private
using namespace std;

static void dump(const string& fn, const void* p, size_t bytes)
{
  FILE* f = fopen(fn.c_str(), "wb");
  if (!f) { perror(fn.c_str()); exit(2); }
  if (bytes) fwrite(p, 1, bytes, f);
  fclose(f);
}
static void slurp(const string& fn, void* p, size_t bytes)
{
  FILE* f = fopen(fn.c_str(), "rb");
  if (!f) { perror(fn.c_str()); exit(2); }
  size_t got = fread(p, 1, bytes, f);
  fclose(f);
  if (got != bytes) { fprintf(stderr, "%s: short read %zu of %zu\n", fn.c_str(), got, bytes); exit(2); }
}

struct AtomLine { string name, species; double x, y, z; };

int main(int argc, char** argv)
{
  if (argc < 3) { fprintf(stderr, "usage: ref_driver basis|run|time <case.txt> [nrep]\n"); return 1; }
  const string mode = argv[1];
  MPI_Init(&argc, &argv);
  if (mode == "xc") {
    // ref_driver xc <prefix>: the reference's own LDAFunctional / PBEFunctional (unpolarized) on the points of
    // <prefix>.in_rho.f64 (n doubles) and <prefix>.in_grad.f64 (3n doubles, component-major); dumps exc, vxc1 (, vxc2)
    const string pre = argv[2];
    FILE* f = fopen((pre + ".in_rho.f64").c_str(), "rb");
    if (!f) { perror("in_rho"); return 2; }
    fseek(f, 0, SEEK_END); const size_t n = ftell(f) / sizeof(double); fclose(f);
    vector<vector<double> > rhoe(1, vector<double>(n));
    vector<double> grad(3 * n);
    slurp(pre + ".in_rho.f64", &rhoe[0][0], n * sizeof(double));
    slurp(pre + ".in_grad.f64", &grad[0], 3 * n * sizeof(double));
    { LDAFunctional lda(rhoe); lda.setxc();
      dump(pre + ".lda_exc.f64", lda.exc, n * sizeof(double)); dump(pre + ".lda_vxc.f64", lda.vxc1, n * sizeof(double)); }
    { PBEFunctional pbe(rhoe);
      for (int j = 0; j < 3; j++) memcpy(pbe.grad_rho[j], &grad[j * n], n * sizeof(double));
      pbe.setxc();
      dump(pre + ".pbe_exc.f64", pbe.exc, n * sizeof(double)); dump(pre + ".pbe_vxc1.f64", pbe.vxc1, n * sizeof(double));
      dump(pre + ".pbe_vxc2.f64", pbe.vxc2, n * sizeof(double)); }
    MPI_Finalize();
    return 0;
  }
  {
  double a[9] = {0}; double ecut = 0, kp[3] = {0,0,0}; int force_complex = 0; int grid[3] = {0,0,0}; int nst = 1;
  vector<pair<string,string> > species; vector<AtomLine> atomlines; string out = "case";
  ifstream in(argv[2]);
  if (!in) { fprintf(stderr, "cannot open %s\n", argv[2]); return 1; }
  string line;
  while (getline(in, line)) {
    istringstream is(line); string key; is >> key;
    if (key == "cell") for (int i = 0; i < 9; i++) is >> a[i];
    else if (key == "ecut") is >> ecut;
    else if (key == "kpoint") is >> kp[0] >> kp[1] >> kp[2];
    else if (key == "force_complex") is >> force_complex;
    else if (key == "grid") is >> grid[0] >> grid[1] >> grid[2];
    else if (key == "nst") is >> nst;
    else if (key == "species") { string n, f; is >> n >> f; species.push_back(make_pair(n, f)); }
    else if (key == "atom") { AtomLine al; is >> al.name >> al.species >> al.x >> al.y >> al.z; atomlines.push_back(al); }
    else if (key == "out") is >> out;
  }

  Context ctxt(1,1);
  Context ctxtsq(ctxt,1,1,0,0);
  Context colctxt(ctxt,1,1,0,0);
  UnitCell cell(D3vector(a[0],a[1],a[2]), D3vector(a[3],a[4],a[5]), D3vector(a[6],a[7],a[8]));
  D3vector kpoint(kp[0],kp[1],kp[2]);

  if (mode == "usx") {
    // ---- SURVEY section 8 row f4, remainder: the ultrasoft branch of NonLocalPotential::energy (D_nm^I from veff(G) and
    //      Q_nm(G), E_nl, H psi; NonLocalPotential.cc:1554-1752) and the augmentation charges of ChargeDensity::update_density
    //      (ChargeDensity.cc:312-465), both run by the reference's own classes on a Sample set up the way SpeciesCmd / RunCmd do
    //      (ultrasoft flag on ctrl and wf, allocation through randomize), with the coefficients and occupations overwritten.
    Sample* s = new Sample(ctxt);
    s->ctrl.ecutden = 0.0; s->ctrl.ultrasoft = true; s->ctrl.nlcc = false; s->ctrl.tddft_involved = false; s->ctrl.extra_memory = 0;
    s->atoms.set_cell(cell);
    for (size_t i = 0; i < species.size(); i++) {
      SpeciesReader rd(ctxt);
      Species* sp = new Species(ctxt, species[i].first);
      rd.readSpecies(*sp, species[i].second);
      rd.bcastSpecies(*sp);
      s->atoms.addSpecies(sp, species[i].first);
      if (!sp->ultrasoft()) { fprintf(stderr, "species %zu is not ultrasoft\n", i); return 3; }
    }
    for (size_t i = 0; i < atomlines.size(); i++)
      s->atoms.addAtom(new Atom(atomlines[i].name, atomlines[i].species, D3vector(atomlines[i].x, atomlines[i].y, atomlines[i].z), D3vector(0,0,0)));
    s->wf.set_ultrasoft(true);
    s->wf.set_cell(cell);
    s->wf.set_ecut(ecut);
    s->wf.set_nel(2*nst);
    s->wf.set_nspin(1);
    if (kp[0] != 0.0 || kp[1] != 0.0 || kp[2] != 0.0) s->wf.add_kpoint(kpoint, 1.0);   // the first added k-point replaces the default one (Wavefunction.cc:1049-1055)
    s->wf.randomize_us(0.01, s->atoms, false);          // allocates, as RunCmd.cc:111-112 does (Wavefunction.cc:1116-1137)
    SlaterDet* sd = s->wf.sd(0,0);
    if (sd->nst() != nst || sd->basis().real()) { fprintf(stderr, "usx: unexpected wavefunction (nst %d, real %d)\n", sd->nst(), (int)sd->basis().real()); return 3; }
    const Basis& basis = sd->basis();
    const int ngw = basis.localsize(), mloc = sd->c().mloc();
    vector<complex<double> > cin((size_t)mloc*nst);
    vector<double> occ(nst);
    slurp(out + ".in_c.f64", &cin[0], cin.size()*sizeof(complex<double>));
    slurp(out + ".in_occ.f64", &occ[0], nst*sizeof(double));
    memcpy(sd->c().valptr(), &cin[0], cin.size()*sizeof(complex<double>));
    sd->set_occ(occ);
    ChargeDensity cd(*s);
    Basis* vb = cd.vbasis();
    const int ngv = vb->localsize();
    const int gr[3] = { cd.vft()->np0(), cd.vft()->np1(), cd.vft()->np2() };
    const size_t N = cd.vft()->np012loc();
    // the density without the augmentation charges: the reference's own SlaterDet::compute_density on the same grid
    vector<double> rho_nc(N, 0.0);
    sd->compute_density(*cd.ft(0,0), 1.0, &rho_nc[0]);
    sd->init_usfns(&s->atoms);
    cd.update_usfns();
    cd.update_density();
    { int hdr[8] = { gr[0], gr[1], gr[2], ngv, vb->nrod_loc(), vb->real() ? 1 : 0, ngw, mloc };
      dump(out + ".usx.hdr.i32", hdr, sizeof hdr);
      vector<int> rods(4*vb->nrod_loc());
      const int nr = vb->nrod_loc();
      for (int i = 0; i < nr; i++) { rods[i] = vb->rod_h(i); rods[nr+i] = vb->rod_k(i); rods[2*nr+i] = vb->rod_lmin(i); rods[3*nr+i] = vb->rod_size(i); }
      dump(out + ".usx.vrods.i32", &rods[0], rods.size()*sizeof(int));
      int mm[2] = { vb->idxmin(1), vb->idxmax(1) };
      dump(out + ".usx.vidxmm.i32", mm, sizeof mm);
      dump(out + ".usx.vkpgx.f64", vb->kpgx_ptr(0), 3*(size_t)ngv*sizeof(double));
      dump(out + ".usx.vg2.f64", vb->g2_ptr(), (size_t)ngv*sizeof(double)); }
    dump(out + ".usx.rho_nc.f64", &rho_nc[0], N*sizeof(double));
    dump(out + ".usx.rho.f64", &cd.rhor[0][0], N*sizeof(double));
    double nel = cd.nelectrons();
    dump(out + ".usx.nel.f64", &nel, sizeof nel);
    // a seeded effective potential on the density basis (an input of NonLocalPotential::energy; EnergyFunctional.cc:924-927)
    vector<complex<double> > veff(ngv);
    { const double* g2 = vb->g2_ptr();
      for (int ig = 0; ig < ngv; ig++) {
        const double ph = 0.37*ig + 1.3*g2[ig];
        veff[ig] = 0.8*exp(-0.15*g2[ig]) * complex<double>(cos(ph), sin(ph));
      } }
    dump(out + ".usx.veff.f64", &veff[0], (size_t)ngv*sizeof(complex<double>));
    NonLocalPotential nlp(s->atoms, sd->context(), basis, 0, false);
    nlp.update_usfns(*sd, vb);
    SlaterDet dsd(*sd); dsd.c().clear();
    vector<vector<double> > fion; valarray<double> sigma(6);
    const double enl = nlp.energy(*sd, true, dsd, false, fion, false, sigma, veff);
    dump(out + ".usx.enl.f64", &enl, sizeof enl);
    dump(out + ".usx.hnl.f64", dsd.c().cvalptr(), (size_t)mloc*nst*sizeof(complex<double>));
    vector<vector<double> > tau; s->atoms.get_positions(tau, true);
    for (int is = 0; is < s->atoms.nsp(); is++) {
      Species* sp = s->atoms.species_list[is];
      char tag[32]; snprintf(tag, sizeof tag, ".usx%d", is);
      const int na = s->atoms.na(is), nlm = sp->nbetalm(), nq = sp->nqtot();
      int h[4] = { na, nlm, nq, 0 };
      dump(out + tag + ".hdr.i32", h, sizeof h);
      vector<int> l(nlm), lm1(nq), lm2(nq);
      vector<double> dzero(nq);
      for (int lm = 0; lm < nlm; lm++) l[lm] = sp->betalm_l(lm);
      for (int qi = 0; qi < nq; qi++) { lm1[qi] = sp->qnm_lm1(qi); lm2[qi] = sp->qnm_lm2(qi); dzero[qi] = sp->dzero(qi); }
      dump(out + tag + ".l.i32", &l[0], nlm*sizeof(int));
      dump(out + tag + ".lm1.i32", &lm1[0], nq*sizeof(int));
      dump(out + tag + ".lm2.i32", &lm2[0], nq*sizeof(int));
      dump(out + tag + ".dzero.f64", &dzero[0], nq*sizeof(double));
      const complex<double>* bg = sd->betag(is)->cvalptr();
      const int bg_mloc = sd->betag(is)->mloc();
      vector<double> tw((size_t)nlm*ngw);
      for (int lm = 0; lm < nlm; lm++) {
        const complex<double> il = l[lm] == 0 ? complex<double>(1,0) : l[lm] == 1 ? complex<double>(0,-1) : l[lm] == 2 ? complex<double>(-1,0) : complex<double>(0,1);
        for (int ig = 0; ig < ngw; ig++) tw[(size_t)lm*ngw + ig] = real(bg[(size_t)lm*bg_mloc + ig] / il);
      }
      dump(out + tag + ".betag.f64", &tw[0], tw.size()*sizeof(double));
      dump(out + tag + ".tau.f64", &tau[is][0], 3*na*sizeof(double));
      // Q_nm(G) on the density basis: the table both NonLocalPotential::update_usfns (:2719) and ChargeDensity::update_usfns (:793) take
      vector<complex<double> > qnm; vector<double> qaug;
      sp->calc_qnmg(vb, qnm, qaug);
      dump(out + tag + ".qnmg.f64", &qnm[0], (size_t)nq*ngv*sizeof(complex<double>));
      const ComplexMatrix* bp = sd->betapsi(is);
      vector<complex<double> > bpo((size_t)nst*na*nlm);
      for (int n = 0; n < nst; n++)
        for (int i = 0; i < na*nlm; i++) bpo[(size_t)n*na*nlm + i] = bp->cvalptr()[(size_t)n*bp->mloc() + i];
      dump(out + tag + ".betapsi.f64", &bpo[0], bpo.size()*sizeof(complex<double>));
    }
    MPI_Finalize();
    return 0;
  }
  SlaterDet sd(ctxt, colctxt, ctxtsq, kpoint, mode == "us", force_complex != 0);   // (ultrasoft forces complex states, SlaterDet.cc:57-58)
  sd.set_nblocks(1,1);
  sd.resize(cell, cell, ecut, nst);
  const Basis& basis = sd.basis();

  if (grid[0] == 0) {
    // ChargeDensity::initialize grid rule (ChargeDensity.cc:77-99): density basis at 4*ecut, +2, factorizable
    Basis vbasis(colctxt, D3vector(0,0,0), false);
    vbasis.resize(cell, cell, 4.0*ecut);
    for (int d = 0; d < 3; d++) { grid[d] = vbasis.np(d) + 2; while (!vbasis.factorizable(grid[d])) grid[d] += 2; }
  }
  FourierTransform ft(basis, grid[0], grid[1], grid[2]);
  const int ngw = basis.localsize();
  const int mloc = sd.c().mloc();
  const int nrods = basis.nrod_loc();
  const size_t N = ft.np012loc();

  {
    int hdr[16] = { grid[0], grid[1], grid[2], ngw, nrods, basis.real() ? 1 : 0, mloc, nst,
                    basis.np(0), basis.np(1), basis.np(2), basis.idxmin(1), basis.idxmax(1), (int)species.size(), 0, 0 };
    dump(out + ".hdr.i32", hdr, sizeof(hdr));
    vector<int> rods(4*nrods);
    for (int i = 0; i < nrods; i++) { rods[i] = basis.rod_h(i); rods[nrods+i] = basis.rod_k(i);
      rods[2*nrods+i] = basis.rod_lmin(i); rods[3*nrods+i] = basis.rod_size(i); }
    dump(out + ".rods.i32", &rods[0], rods.size()*sizeof(int));
    dump(out + ".idx.i32", basis.idx_ptr(), 3*ngw*sizeof(int));
    dump(out + ".kpg2.f64", basis.kpg2_ptr(), ngw*sizeof(double));
    dump(out + ".kpgx.f64", basis.kpgx_ptr(0), 3*ngw*sizeof(double));
    double om = cell.volume();
    dump(out + ".omega.f64", &om, sizeof(double));
  }
  if (mode == "basis") { MPI_Finalize(); return 0; }

  // ---- atoms / species / projector tables
  AtomSet atoms(ctxt);
  atoms.set_cell(cell);
  for (size_t i = 0; i < species.size(); i++) {
    SpeciesReader rd(ctxt);
    Species* sp = new Species(ctxt, species[i].first);
    rd.readSpecies(*sp, species[i].second);
    rd.bcastSpecies(*sp);
    atoms.addSpecies(sp, species[i].first);
  }
  for (size_t i = 0; i < atomlines.size(); i++) {
    Atom* at = new Atom(atomlines[i].name, atomlines[i].species,
                        D3vector(atomlines[i].x, atomlines[i].y, atomlines[i].z), D3vector(0,0,0));
    atoms.addAtom(at);
  }
  if (mode == "us") {
    // ---- the ultrasoft beta.psi path.  SlaterDet::init_usfns (SlaterDet.cc:103-197) runs calc_betag
    //      (:2006-2127), calc_betapsi (:2130-2263), Species::calc_qnmg -> set_qaug, calc_spsi (:2426-2570) on the given states
    vector<complex<double> > cin((size_t)mloc*nst);
    slurp(out + ".in_c.f64", &cin[0], cin.size()*sizeof(complex<double>));
    memcpy(sd.c().valptr(), &cin[0], cin.size()*sizeof(complex<double>));
    sd.init_usfns(&atoms);
    vector<vector<double> > tau; atoms.get_positions(tau, true);
    for (int is = 0; is < atoms.nsp(); is++) {
      Species* s = atoms.species_list[is];
      if (!s->ultrasoft()) { fprintf(stderr, "species %d is not ultrasoft\n", is); return 3; }
      char tag[32]; snprintf(tag, sizeof tag, ".us%d", is);
      const int na = atoms.na(is), nlm = s->nbetalm(), nq = s->nqtot();
      int h[4] = { na, nlm, nq, 0 };
      dump(out + tag + ".hdr.i32", h, sizeof h);
      vector<int> l(nlm), lm1(nq), lm2(nq);
      for (int lm = 0; lm < nlm; lm++) l[lm] = s->betalm_l(lm);
      for (int qi = 0; qi < nq; qi++) { lm1[qi] = s->qnm_lm1(qi); lm2[qi] = s->qnm_lm2(qi); }
      dump(out + tag + ".l.i32", &l[0], nlm*sizeof(int));
      dump(out + tag + ".lm1.i32", &lm1[0], nq*sizeof(int));
      dump(out + tag + ".lm2.i32", &lm2[0], nq*sizeof(int));
      // betag without the structure factor (the !highmem branch): bval * ylm * (-i)^l -> the real table bval * ylm
      const complex<double>* bg = sd.betag(is)->cvalptr();
      const int bg_mloc = sd.betag(is)->mloc();
      vector<double> tw((size_t)nlm*ngw);
      for (int lm = 0; lm < nlm; lm++) {
        const complex<double> il = l[lm] == 0 ? complex<double>(1,0) : l[lm] == 1 ? complex<double>(0,-1) : l[lm] == 2 ? complex<double>(-1,0) : complex<double>(0,1);
        for (int ig = 0; ig < ngw; ig++) tw[(size_t)lm*ngw + ig] = real(bg[(size_t)lm*bg_mloc + ig] / il);
      }
      dump(out + tag + ".betag.f64", &tw[0], tw.size()*sizeof(double));
      dump(out + tag + ".tau.f64", &tau[is][0], 3*na*sizeof(double));
      vector<complex<double> > qnm; vector<double> qaug;
      s->calc_qnmg(const_cast<Basis*>(&basis), qnm, qaug);
      dump(out + tag + ".qaug.f64", &qaug[0], nq*sizeof(double));
      // betapsi: (na*nlm) x nst, element (ia*nlm + lm, n)
      const ComplexMatrix* bp = sd.betapsi(is);
      vector<complex<double> > bpo((size_t)nst*na*nlm);
      for (int n = 0; n < nst; n++)
        for (int i = 0; i < na*nlm; i++) bpo[(size_t)n*na*nlm + i] = bp->cvalptr()[(size_t)n*bp->mloc() + i];
      dump(out + tag + ".betapsi.f64", &bpo[0], bpo.size()*sizeof(complex<double>));
    }
    dump(out + ".spsi.f64", sd.spsi().cvalptr(), (size_t)mloc*nst*sizeof(complex<double>));
    MPI_Finalize();
    return 0;
  }
  NonLocalPotential* nlp = 0;
  if (!species.empty()) {
    nlp = new NonLocalPotential(atoms, ctxt, basis, 0, false);
    nlp->update_twnl(false);   // as EnergyFunctional does after construction (EnergyFunctional.cc:2036)
    vector<vector<double> > tau; atoms.get_positions(tau, true);
    for (int is = 0; is < nlp->nsp; is++) {
      char tag[32]; snprintf(tag, sizeof tag, ".sp%d", is);
      int h[4] = { nlp->na[is], nlp->npr[is], 0, 0 };
      dump(out + tag + ".hdr.i32", h, sizeof h);
      dump(out + tag + ".lproj.i32", nlp->npr[is] ? &nlp->lproj[is][0] : 0, nlp->npr[is]*sizeof(int));
      dump(out + tag + ".wt.f64", nlp->npr[is] ? &nlp->wt[is][0] : 0, nlp->npr[is]*sizeof(double));
      dump(out + tag + ".twnl.f64", nlp->npr[is] ? &nlp->twnl[is][0] : 0, (size_t)nlp->npr[is]*ngw*sizeof(double));
      dump(out + tag + ".tau.f64", &tau[is][0], 3*nlp->na[is]*sizeof(double));
      // what NonLocalPotential::update_twnl (NonLocalPotential.cc:261-1522) builds twnl FROM, for
      // Kleinman-Bylander species (nquad == 0): per projector its m and its radial table (l, channel), the tables being the
      // species' cubic splines behind Species::dvnlg (Species.cc:1492-1505; spline.cc:126-156)
      Species* s = atoms.species_list[is];
      int kb[4] = { nlp->nquad[is], s->ndft_, 0, 0 };
      if (nlp->npr[is] > 0 && nlp->nquad[is] == 0) {
        vector<int> mproj(nlp->npr[is], -1), tproj(nlp->npr[is], -1);
        vector<double> ytab, y2tab;
        int ntab = 0;
        for (int l = 0; l <= nlp->lmax[is]; l++) {
          if (l == nlp->lloc[is]) continue;
          for (int ic = 0; ic < s->nchannels(); ic++) {
            const Spline& sp = s->projectors_g_[l][ic];
            ytab.insert(ytab.end(), sp.y_.begin(), sp.y_.end());
            y2tab.insert(y2tab.end(), sp.y2_.begin(), sp.y2_.end());
            for (int m = 0; m < 2*l+1; m++) { const int ipr = nlp->iprojlm[is][l][m][ic]; mproj[ipr] = m; tproj[ipr] = ntab; }
            ntab++;
          }
        }
        kb[2] = ntab;
        dump(out + tag + ".kb_m.i32", &mproj[0], mproj.size()*sizeof(int));
        dump(out + tag + ".kb_tab.i32", &tproj[0], tproj.size()*sizeof(int));
        dump(out + tag + ".kb_gspl.f64", &s->gspl_[0], s->ndft_*sizeof(double));
        dump(out + tag + ".kb_y.f64", &ytab[0], ytab.size()*sizeof(double));
        dump(out + tag + ".kb_y2.f64", &y2tab[0], y2tab.size()*sizeof(double));
      }
      if (nlp->npr[is] > 0 && nlp->nquad[is] > 0) {
        // semi-local species (nquad > 0): projector ipr = iquad + nquad * ilm (NonLocalPotential.cc:234-249); per projector its m
        // and the quadrature radius its radial function 4 pi j_l(|k+G| r) r is taken at
        vector<int> mproj(nlp->npr[is], -1);
        vector<double> rproj(nlp->npr[is], 0.0);
        int ilm = 0;
        for (int l = 0; l <= nlp->lmax[is]; l++) {
          if (l == nlp->lloc[is]) continue;
          for (int m = 0; m < 2*l+1; m++, ilm++)
            for (int iq = 0; iq < nlp->nquad[is]; iq++) { const int ipr = iq + nlp->nquad[is]*ilm; mproj[ipr] = m; rproj[ipr] = nlp->rquad[is][iq]; }
        }
        dump(out + tag + ".sl_m.i32", &mproj[0], mproj.size()*sizeof(int));
        dump(out + tag + ".sl_r.f64", &rproj[0], rproj.size()*sizeof(double));
      }
      dump(out + tag + ".kb.i32", kb, sizeof kb);
    }
  }

  // ---- inputs
  vector<complex<double> > cin((size_t)mloc*nst);
  vector<double> v(N), occ(nst);
  slurp(out + ".in_c.f64", &cin[0], cin.size()*sizeof(complex<double>));
  slurp(out + ".in_v.f64", &v[0], N*sizeof(double));
  slurp(out + ".in_occ.f64", &occ[0], nst*sizeof(double));
  memcpy(sd.c().valptr(), &cin[0], cin.size()*sizeof(complex<double>));
  sd.set_occ(occ);

  if (mode == "time") {
    // CPU baseline: the reference's own per-state loops, wall-clocked (used by bench.py's cpu_baseline leg)
    const int nrep = argc > 3 ? atoi(argv[3]) : 1;
    SlaterDet dsd(sd); dsd.c().clear();
    vector<double> rho(N, 0.0);
    vector<vector<double> > fion; valarray<double> sigma(6); vector<complex<double> > veff;
    double t_loc = 0, t_nl = 0, t_rho = 0, t_kin = 0;
    for (int r = 0; r < nrep; r++) {
      Timer t0; t0.start(); if (nlp) nlp->energy(sd, true, dsd, false, fion, false, sigma, veff); t0.stop(); t_nl += t0.real();
      Timer t1; t1.start();
      { const double* kpg2 = basis.kpg2_ptr(); complex<double>* cp = dsd.c().valptr(); const complex<double>* c = sd.c().cvalptr();
        for (int n = 0; n < sd.nstloc(); n++) for (int ig = 0; ig < ngw; ig++) cp[ig+mloc*n] += 0.5 * kpg2[ig] * c[ig+mloc*n]; }
      t1.stop(); t_kin += t1.real();
      Timer t2; t2.start(); sd.rs_mul_add(ft, &v[0], dsd); t2.stop(); t_loc += t2.real();
      Timer t3; t3.start(); sd.compute_density(ft, 1.0, &rho[0]); t3.stop(); t_rho += t3.real();
    }
    printf("{\"nst\": %d, \"nrep\": %d, \"threads\": %d, \"t_nonlocal\": %.6f, \"t_kinetic\": %.6f, \"t_local\": %.6f, \"t_density\": %.6f}\n",
           nst, nrep, omp_get_max_threads(), t_nl, t_kin, t_loc, t_rho);
    MPI_Finalize(); return 0;
  }

  // ---- single transforms on state 0: f = backward(c0); g = v*f; c' = forward(g)
  {
    vector<complex<double> > f(N), cc(mloc);
    ft.backward(sd.c().cvalptr(0), &f[0]);
    dump(out + ".bwd0.f64", &f[0], N*sizeof(complex<double>));
    for (size_t i = 0; i < N; i++) f[i] *= v[i];
    ft.forward(&f[0], &cc[0]);
    dump(out + ".fwd0.f64", &cc[0], ngw*sizeof(complex<double>));
    if (basis.real() && nst >= 2) {
      vector<complex<double> > c1(mloc), c2(mloc);
      ft.backward(sd.c().cvalptr(0), sd.c().cvalptr(mloc), &f[0]);
      dump(out + ".bwdpair01.f64", &f[0], N*sizeof(complex<double>));
      for (size_t i = 0; i < N; i++) f[i] *= v[i];
      ft.forward(&f[0], &c1[0], &c2[0]);
      dump(out + ".fwdpair0.f64", &c1[0], ngw*sizeof(complex<double>));
      dump(out + ".fwdpair1.f64", &c2[0], ngw*sizeof(complex<double>));
    }
  }

  // ---- rs_mul_add alone (sdp starts at zero)
  {
    SlaterDet dsd(sd); dsd.c().clear();
    sd.rs_mul_add(ft, &v[0], dsd);
    dump(out + ".hloc.f64", dsd.c().cvalptr(), (size_t)mloc*nst*sizeof(complex<double>));
  }
  // ---- density
  {
    vector<double> rho(N, 0.0);
    sd.compute_density(ft, 1.0, &rho[0]);
    dump(out + ".rho.f64", &rho[0], N*sizeof(double));
  }
  // ---- nonlocal alone, then the whole H psi in the reference's order: clear -> nonlocal -> kinetic -> local
  {
    SlaterDet dsd(sd); dsd.c().clear();
    vector<vector<double> > fion; valarray<double> sigma(6); vector<complex<double> > veff;
    double enl = 0.0;
    if (nlp) {
      enl = nlp->energy(sd, true, dsd, false, fion, false, sigma, veff);
      dump(out + ".hnl.f64", dsd.c().cvalptr(), (size_t)mloc*nst*sizeof(complex<double>));
    }
    dump(out + ".enl.f64", &enl, sizeof(double));
    const double* kpg2 = basis.kpg2_ptr();
    complex<double>* cp = dsd.c().valptr(); const complex<double>* c = sd.c().cvalptr();
    for (int n = 0; n < sd.nstloc(); n++)
      for (int ig = 0; ig < ngw; ig++)
        cp[ig+mloc*n] += 0.5 * kpg2[ig] * c[ig+mloc*n];       // EnergyFunctional.cc:1675-1677
    sd.rs_mul_add(ft, &v[0], dsd);                               // EnergyFunctional.cc:1695
    dump(out + ".hpsi.f64", dsd.c().cvalptr(), (size_t)mloc*nst*sizeof(complex<double>));
    // ---- descent direction of the PSD/PSDA steppers on (psi, H psi): the reference's own matrix calls, in the order of
    //      PSDAWavefunctionStepper::update (PSDAWavefunctionStepper.cc:65-84 real, :264-277 complex)
    if (basis.real()) {
      DoubleMatrix c_proxy(sd.c());
      DoubleMatrix cp_proxy(dsd.c());
      DoubleMatrix am(c_proxy.context(), c_proxy.n(), c_proxy.n(), c_proxy.nb(), c_proxy.nb());
      am.gemm('t','n',2.0,c_proxy,cp_proxy,0.0);
      am.ger(-1.0,c_proxy,0,cp_proxy,0);
      cp_proxy.gemm('n','n',-1.0,c_proxy,am,1.0);
      dump(out + ".resid_a.f64", am.cvalptr(), (size_t)nst*nst*sizeof(double));
      { // Wavefunction::diag (Wavefunction.cc:1538-1539, 1612): eigenvalues of the same h through the reference's syevd('l')
        DoubleMatrix hd(am); valarray<double> w(hd.m()); hd.syevd('l', w);
        dump(out + ".diag_w.f64", &w[0], nst*sizeof(double)); }
    } else {
      ComplexMatrix& c_proxy = sd.c();
      ComplexMatrix& cpm = dsd.c();
      ComplexMatrix am(c_proxy.context(), c_proxy.n(), c_proxy.n(), c_proxy.nb(), c_proxy.nb());
      am.gemm('c','n',1.0,c_proxy,cpm,0.0);
      cpm.gemm('n','n',-1.0,c_proxy,am,1.0);
      dump(out + ".resid_a.f64", am.cvalptr(), (size_t)nst*nst*sizeof(complex<double>));
      { // Wavefunction::diag (Wavefunction.cc:1641, 1693): eigenvalues of the same h through the reference's heev('l')
        ComplexMatrix hd(am); valarray<double> w(hd.m()); hd.heev('l', w);
        dump(out + ".diag_w.f64", &w[0], nst*sizeof(double)); }
    }
    dump(out + ".resid.f64", dsd.c().cvalptr(), (size_t)mloc*nst*sizeof(complex<double>));
  }
  // ---- current density: SlaterDet::compute_density(ft, weight, complex* rho, sd2) (SlaterDet.cc:935-968) driven as
  //      CurrentDensity::update_current does (CurrentDensity.cc:52-86; that class needs a whole Sample, its loop is restated)
  {
    vector<double> cur(3*N, 0.0);
    vector<complex<double> > tmp(N);
    SlaterDet rsd(sd);
    for (int idir = 0; idir < 3; idir++) {
      const double* kx = basis.kpgx_ptr(idir);
      for (int n = 0; n < sd.nstloc(); n++)
        for (int ig = 0; ig < ngw; ig++)
          rsd.c()[ig + mloc*n] = complex<double>(0.0, 1.0) * kx[ig] * sd.c()[ig + mloc*n];   // CurrentDensity.cc:72-76
      for (size_t i = 0; i < N; i++) tmp[i] = 0.0;
      sd.compute_density(ft, 1.0, &tmp[0], rsd);
      for (size_t i = 0; i < N; i++) cur[idir*N + i] += -imag(tmp[i]);                        // CurrentDensity.cc:85-87
    }
    dump(out + ".cur.f64", &cur[0], 3*N*sizeof(double));
  }
  // ---- SlaterDet::gram (SlaterDet.cc:1043-1143) on the input coefficients
  {
    SlaterDet gsd(sd);
    gsd.gram();
    dump(out + ".gram.f64", gsd.c().cvalptr(), (size_t)mloc*nst*sizeof(complex<double>));
  }
  delete nlp;
  }
  MPI_Finalize();
  return 0;
}
