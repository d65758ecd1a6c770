#!/usr/bin/env python
"""integration/apply_shim.py -- writes USE_QB200 copies of the reference's four seam sources.

    python integration/apply_shim.py <reference-root> <out-dir>

Reads src/qball/{FourierTransform,SlaterDet,NonLocalPotential,EnergyFunctional}.cc where they lie under the (read-only)
reference tree and writes patched copies into <out-dir> (a build directory: git-ignored, nothing of the reference is
committed).  Each patch is a few forwarding lines inserted at the top of a member function -- the reference's signatures,
headers and every other source stay untouched (INTEGRATION.md sections 1-5).  Insertion points are found by the function
signatures, not by line numbers, and every one must match exactly once.
"""
import os
import re
import sys


def insert_after_open_brace(src, signature_regex, code, what):
    """insert `code` right after the first '{' that follows the (single) match of signature_regex"""
    ms = list(re.finditer(signature_regex, src, flags=re.M))
    if len(ms) != 1:
        raise SystemExit(f"apply_shim: {what}: expected exactly one match, found {len(ms)}")
    brace = src.index("{", ms[0].end())
    return src[:brace + 1] + "\n" + code + src[brace + 1:]


def insert_before(src, anchor_regex, code, what, after_regex=None):
    """insert `code` before the first match of anchor_regex located after the (single) match of after_regex"""
    start = 0
    if after_regex:
        ms = list(re.finditer(after_regex, src, flags=re.M))
        if len(ms) != 1:
            raise SystemExit(f"apply_shim: {what}: context: expected exactly one match, found {len(ms)}")
        start = ms[0].end()
    m = re.compile(anchor_regex, flags=re.M).search(src, start)
    if not m:
        raise SystemExit(f"apply_shim: {what}: anchor not found")
    return src[:m.start()] + code + src[m.start():]


INCLUDE = '#ifdef USE_QB200\n#include "qb200_shim.h"\n#include <qball_b200.hpp>\n#endif\n'


def patch_fourier_transform(s):
    s = INCLUDE + s
    s = insert_after_open_brace(s, r"^FourierTransform::~FourierTransform\(\)", "#ifdef USE_QB200\n  qb200_shim::ft_detach(this);\n#endif\n", "~FourierTransform")
    s = insert_after_open_brace(s, r"^FourierTransform::FourierTransform \(const Basis &basis,\s*\n\s*int np0, int np1, int np2\)[^{]*",
                                "#ifdef USE_QB200\n  qb200_shim::ft_attach(this, basis, np0, np1, np2);\n#endif\n", "FourierTransform ctor")
    fwd = "#ifdef USE_QB200\n  if (qb200_shim::enabled()) { qb200_shim::ft_gpu(this).%s; return; }\n#endif\n"
    s = insert_after_open_brace(s, r"^void FourierTransform::backward\(const complex<double>\* c, complex<double>\* f\)", fwd % "backward(c, f)", "backward")
    s = insert_after_open_brace(s, r"^void FourierTransform::forward\(complex<double>\* f, complex<double>\* c\)", fwd % "forward(f, c)", "forward")
    s = insert_after_open_brace(s, r"^void FourierTransform::backward\(const complex<double>\* c1,\s*\n\s*const complex<double>\* c2,\s*\n\s*complex<double>\* f\)",
                                fwd % "backward(c1, c2, f)", "backward pair")
    s = insert_after_open_brace(s, r"^void FourierTransform::forward\(complex<double>\* f,\s*\n\s*complex<double>\* c1, complex<double>\* c2\)",
                                fwd % "forward(f, c1, c2)", "forward pair")
    return s


def patch_slater_det(s):
    s = INCLUDE + s
    s = insert_after_open_brace(s, r"^void SlaterDet::compute_density\(FourierTransform& ft,\s*\n\s*double weight, double\* rho\) const",
                                "#ifdef USE_QB200\n  if (qb200_shim::enabled()) {\n    qb200_shim::compute_density(ft, c_, nstloc(), occ_, weight, basis_->cell().volume(), rho);\n    return;\n  }\n#endif\n",
                                "SlaterDet::compute_density")
    s = insert_after_open_brace(s, r"^void SlaterDet::rs_mul_add\(FourierTransform& ft,\s*\n\s*const double\* v, SlaterDet& sdp\) const",
                                "#ifdef USE_QB200\n  if (qb200_shim::enabled()) { qb200_shim::rs_mul_add(ft, c_, nstloc(), v, sdp.c()); return; }\n#endif\n",
                                "SlaterDet::rs_mul_add")
    return s


def patch_nonlocal(s):
    s = INCLUDE + s
    s = insert_after_open_brace(s, r"^NonLocalPotential::~NonLocalPotential\(void\)", "#ifdef USE_QB200\n  qb200_shim::nl_invalidate(this);\n#endif\n", "~NonLocalPotential")
    s = insert_after_open_brace(s, r"^void NonLocalPotential::update_twnl\(const bool compute_stress\)",
                                "#ifdef USE_QB200\n  qb200_shim::nl_invalidate(this);      // the device tables are rebuilt from the new twnl on the next energy()\n#endif\n",
                                "update_twnl")
    code = ("#ifdef USE_QB200\n"
            "  // norm-conserving branch without forces / stress / vector potential: the whole species loop (:1909-2171) on the device\n"
            "  if (qb200_shim::enabled() && !ultrasoft_ && !compute_forces && !compute_stress && !vp && nspnl > 0) {\n"
            "    double enl_gpu = qb200_shim::nl_energy(this, basis_, atoms_, nsp, na, npr, lproj, wt, twnl, sd, compute_hpsi, dsd);\n"
            "    ctxt_.dsum('r',1,1,&enl_gpu,1);\n"
            "    sigma_enl = 0.0;\n"
            "    return enl_gpu;\n"
            "  }\n"
            "#endif\n")
    s = insert_after_open_brace(s, r"^double NonLocalPotential::energy\(SlaterDet& sd, bool compute_hpsi, SlaterDet& dsd,[^{]*", code, "NonLocalPotential::energy")
    return s


def patch_energy_functional(s):
    s = INCLUDE + s
    # the psi2sum loop: `if (device filled psi2sum) {} else for (...)` -- the reference's loop stays the else branch
    code = ("#ifdef USE_QB200\n"
            "              if (!vp && qb200_shim::psi2sum(ft[ispin][ikp], c, occ, fac, kpg2, psi2sum)) {} else\n"
            "#endif\n")
    s = insert_before(s, r"^[ \t]*for \( int lj=0; lj < c\.nblocks\(\); lj\+\+ \)", code, "EnergyFunctional psi2sum loop",
                      after_regex=r"compute psi2sum\(G\) = fac \* sum_G occ\(n\) psi2\(n,G\)")
    # update_vhxc: one spin, LDA / PBE, none of the special branches -> the whole function on the device
    code = ("#ifdef USE_QB200\n"
            "  if (qb200_shim::enabled() && wf_.nspin() == 1 && !s_.ctrl.ultrasoft && !s_.ctrl.nlcc && s_.ctrl.esm_bc == \"\" &&\n"
            "      s_.ctrl.enthalpy_pressure == 0.0 && !s_.ctrl.tddft_involved && s_.ctrl.vdw != \"D3\" && s_.ctrl.stress != \"ON\" &&\n"
            "      (s_.ctrl.xc == \"LDA\" || s_.ctrl.xc == \"PBE\")) {\n"
            "    const int ngloc_q = vbasis_->localsize();\n"
            "    const double omega_q = wf_.cell().volume();\n"
            "    double en_q[3];\n"
            "    qb200_shim::update_vhxc(vft, s_.ctrl.xc == \"PBE\" ? 1 : 0, &cd_.rhor[0][0], &cd_.rhog[0][0], vbasis_->gx_ptr(0), vbasis_->g2i_ptr(),\n"
            "                            &vion_local_g[0], &rhopst[0], omega_q, &v_r[0][0], &rhogt[0], en_q);\n"
            "    const double *const g2i_q = vbasis_->g2i_ptr();\n"
            "    for ( int ig = 0; ig < ngloc_q; ig++ ) {          // members other parts of energy() read (forces: :1715-1722)\n"
            "      rhoelg[ig] = cd_.rhog[0][ig] / omega_q;\n"
            "      vlocal_g[ig] = vion_local_g[ig] + 4.0 * M_PI * rhogt[ig] * g2i_q[ig];\n"
            "    }\n"
            "    exc_ = en_q[0]; eps_ = en_q[1]; ehart_ = en_q[2];\n"
            "    epv_ = 0.0; evdw_ = 0.0; sigma_vdw = 0.0;\n"
            "    return;\n"
            "  }\n"
            "#endif\n")
    s = insert_after_open_brace(s, r"^void EnergyFunctional::update_vhxc\(void\)", code, "EnergyFunctional::update_vhxc")
    return s


PATCHES = {"FourierTransform.cc": patch_fourier_transform, "SlaterDet.cc": patch_slater_det,
           "NonLocalPotential.cc": patch_nonlocal, "EnergyFunctional.cc": patch_energy_functional}


def main():
    ref, out = sys.argv[1], sys.argv[2]
    os.makedirs(out, exist_ok=True)
    for name, fn in PATCHES.items():
        src = open(os.path.join(ref, "src", "qball", name), encoding="latin-1").read()
        patched = fn(src)
        with open(os.path.join(out, name), "w", encoding="latin-1") as f:
            f.write(patched)
        print(f"apply_shim: {name}: +{patched.count(chr(10)) - src.count(chr(10))} lines")


if __name__ == "__main__":
    main()
