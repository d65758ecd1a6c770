"""oracle/refdrive.py -- TEST INFRASTRUCTURE ONLY (never imported by the product package).

Drives ``oracle/_ref/ref_driver`` (the UNMODIFIED reference, compiled from /root/reference by ``oracle/Makefile``)
and defines the deterministic synthetic inputs used by the golden fixtures, the parity tests
and bench.py.  The generator is counter based (splitmix64) so any slice can be regenerated without storing it.
"""
from __future__ import annotations

import os
import subprocess
import tempfile
from dataclasses import dataclass, field

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DRIVER = os.path.join(HERE, "_ref", "ref_driver")


def have_ref() -> bool:
    return os.path.isfile(REF_DRIVER) and os.access(REF_DRIVER, os.X_OK)


# --------------------------------------------------------------------------------------------- synthetic inputs
# one generator for fixtures, tests and bench: it lives in the package (pure numpy input synthesis, no path arithmetic)
import sys as _sys
_sys.path.insert(0, os.path.dirname(HERE))
from qball_b200.synth import splitmix_uniform, synth_coefficients, synth_occ, synth_potential  # noqa: E402,F401


# --------------------------------------------------------------------------------------------- reference driver
@dataclass
class Case:
    cell: tuple  # 9 numbers
    ecut: float  # hartree
    kpoint: tuple = (0.0, 0.0, 0.0)
    force_complex: bool = False
    grid: tuple = (0, 0, 0)
    nst: int = 2
    species: list = field(default_factory=list)  # [(name, xml path)]
    atoms: list = field(default_factory=list)  # [(name, species, x, y, z)]

    def text(self, out_prefix: str) -> str:
        lines = ["cell " + " ".join(repr(float(x)) for x in self.cell), f"ecut {self.ecut!r}",
                 "kpoint " + " ".join(repr(float(x)) for x in self.kpoint),
                 f"force_complex {int(self.force_complex)}", "grid " + " ".join(str(int(g)) for g in self.grid),
                 f"nst {self.nst}"]
        lines += [f"species {n} {p}" for n, p in self.species]
        lines += [f"atom {n} {s} {x!r} {y!r} {z!r}" for n, s, x, y, z in self.atoms]
        lines.append(f"out {out_prefix}")
        return "\n".join(lines) + "\n"


def _load(prefix, name, dtype):
    return np.fromfile(f"{prefix}.{name}", dtype=dtype)


def read_basis(prefix: str) -> dict:
    hdr = _load(prefix, "hdr.i32", np.int32)
    b = dict(np0=int(hdr[0]), np1=int(hdr[1]), np2=int(hdr[2]), ngw=int(hdr[3]), nrods=int(hdr[4]),
             is_real=bool(hdr[5]), mloc=int(hdr[6]), nst=int(hdr[7]), basis_np=tuple(int(x) for x in hdr[8:11]),
             idxmin1=int(hdr[11]), idxmax1=int(hdr[12]), nsp=int(hdr[13]))
    rods = _load(prefix, "rods.i32", np.int32).reshape(4, b["nrods"])
    b.update(rod_h=rods[0].copy(), rod_k=rods[1].copy(), rod_lmin=rods[2].copy(), rod_size=rods[3].copy())
    b["idx"] = _load(prefix, "idx.i32", np.int32).reshape(b["ngw"], 3)
    b["kpg2"] = _load(prefix, "kpg2.f64", np.float64)
    b["kpgx"] = _load(prefix, "kpgx.f64", np.float64).reshape(3, b["ngw"])
    b["omega"] = float(_load(prefix, "omega.f64", np.float64)[0])
    return b


def run_reference(case: Case, seed: int = 1, nocc: int | None = None, workdir: str | None = None,
                  inputs: dict | None = None, threads: int = 8) -> dict:
    """basis phase -> synthetic (or given) inputs -> run phase; returns every dumped array."""
    assert have_ref(), "oracle/_ref/ref_driver missing: run `make -C oracle ref` where /root/reference exists"
    tmp = workdir or tempfile.mkdtemp(prefix="qbref_")
    prefix = os.path.join(tmp, "case")
    cf = os.path.join(tmp, "case.txt")
    with open(cf, "w") as f:
        f.write(case.text(prefix))
    env = dict(os.environ, OMP_NUM_THREADS=str(threads))
    subprocess.run([REF_DRIVER, "basis", cf], check=True, env=env, stdout=subprocess.DEVNULL)
    b = read_basis(prefix)
    if inputs is None:
        c = synth_coefficients(b["kpg2"], case.ecut, case.nst, b["mloc"], b["is_real"], seed)
        v = synth_potential(b["np0"], b["np1"], b["np2"], seed + 6)
        occ = synth_occ(case.nst, nocc)
    else:
        c, v, occ = inputs["c"], inputs["v"], inputs["occ"]
    c.tofile(prefix + ".in_c.f64")
    v.tofile(prefix + ".in_v.f64")
    occ.tofile(prefix + ".in_occ.f64")
    subprocess.run([REF_DRIVER, "run", cf], check=True, env=env, stdout=subprocess.DEVNULL)
    N = b["np0"] * b["np1"] * b["np2"]
    r = dict(b)
    r.update(c=c, v=v, occ=occ)
    r["bwd0"] = _load(prefix, "bwd0.f64", np.complex128)
    r["fwd0"] = _load(prefix, "fwd0.f64", np.complex128)
    if b["is_real"] and case.nst >= 2:
        r["bwdpair01"] = _load(prefix, "bwdpair01.f64", np.complex128)
        r["fwdpair0"] = _load(prefix, "fwdpair0.f64", np.complex128)
        r["fwdpair1"] = _load(prefix, "fwdpair1.f64", np.complex128)
    r["hloc"] = _load(prefix, "hloc.f64", np.complex128).reshape(case.nst, b["mloc"])
    r["rho"] = _load(prefix, "rho.f64", np.float64)
    r["hpsi"] = _load(prefix, "hpsi.f64", np.complex128).reshape(case.nst, b["mloc"])
    r["enl"] = float(_load(prefix, "enl.f64", np.float64)[0])
    r["resid"] = _load(prefix, "resid.f64", np.complex128).reshape(case.nst, b["mloc"])
    r["resid_a"] = _load(prefix, "resid_a.f64", np.float64 if b["is_real"] else np.complex128).reshape(case.nst, case.nst)
    if os.path.exists(prefix + ".diag_w.f64"):
        r["diag_w"] = _load(prefix, "diag_w.f64", np.float64)
    r["cur"] = _load(prefix, "cur.f64", np.float64).reshape(3, N)
    r["gram"] = _load(prefix, "gram.f64", np.complex128).reshape(case.nst, b["mloc"])
    sp = []
    for i in range(b["nsp"]):
        h = _load(prefix, f"sp{i}.hdr.i32", np.int32)
        na, npr = int(h[0]), int(h[1])
        sp.append(dict(na=na, npr=npr, lproj=_load(prefix, f"sp{i}.lproj.i32", np.int32),
                       wt=_load(prefix, f"sp{i}.wt.f64", np.float64),
                       twnl=_load(prefix, f"sp{i}.twnl.f64", np.float64).reshape(npr, b["ngw"]),
                       tau=_load(prefix, f"sp{i}.tau.f64", np.float64).reshape(na, 3)))
        kb = _load(prefix, f"sp{i}.kb.i32", np.int32)
        sp[-1]["nquad"] = int(kb[0])
        if npr > 0 and int(kb[0]) > 0:       # semi-local: per projector m and the quadrature radius (row a11)
            sp[-1].update(mproj=_load(prefix, f"sp{i}.sl_m.i32", np.int32), rproj=_load(prefix, f"sp{i}.sl_r.f64", np.float64))
        if npr > 0 and int(kb[0]) == 0:      # Kleinman-Bylander: the radial spline tables update_twnl builds twnl from (row a11)
            ndft, ntab = int(kb[1]), int(kb[2])
            sp[-1].update(mproj=_load(prefix, f"sp{i}.kb_m.i32", np.int32), tabproj=_load(prefix, f"sp{i}.kb_tab.i32", np.int32),
                          gspl=_load(prefix, f"sp{i}.kb_gspl.f64", np.float64),
                          vnlg=_load(prefix, f"sp{i}.kb_y.f64", np.float64).reshape(ntab, ndft),
                          vnlg_spl=_load(prefix, f"sp{i}.kb_y2.f64", np.float64).reshape(ntab, ndft))
    r["species"] = sp
    if b["nsp"]:
        r["hnl"] = _load(prefix, "hnl.f64", np.complex128).reshape(case.nst, b["mloc"])
    assert r["bwd0"].shape[0] == N
    return r


def run_reference_us(case: Case, seed: int = 1, workdir: str | None = None, threads: int = 8) -> dict:
    """SURVEY section 8 row f4 (ultrasoft beta.psi path): ``ref_driver us`` = the reference's own SlaterDet::init_usfns
    (calc_betag, calc_betapsi, Species::calc_qnmg -> set_qaug, calc_spsi) on seeded coefficients.  Returns the basis, the
    coefficients, per species the betag tables (without (-i)^l and without the structure factor), l per channel, the
    (lm1, lm2, qaug) triples, positions and betapsi; and spsi."""
    assert have_ref(), "oracle/_ref/ref_driver missing: run `make -C oracle ref` where /root/reference exists"
    tmp = workdir or tempfile.mkdtemp(prefix="qbrefus_")
    prefix = os.path.join(tmp, "case")
    cf = os.path.join(tmp, "case.txt")
    with open(cf, "w") as f:
        f.write(case.text(prefix))
    env = dict(os.environ, OMP_NUM_THREADS=str(threads))
    subprocess.run([REF_DRIVER, "basis", cf], check=True, env=env, stdout=subprocess.DEVNULL)
    b = read_basis(prefix)
    c = synth_coefficients(b["kpg2"], case.ecut, case.nst, b["mloc"], b["is_real"], seed)
    c.tofile(prefix + ".in_c.f64")
    subprocess.run([REF_DRIVER, "us", cf], check=True, env=env, stdout=subprocess.DEVNULL)
    r = dict(b)
    r["c"] = c
    sp = []
    for i in range(b["nsp"]):
        h = _load(prefix, f"us{i}.hdr.i32", np.int32)
        na, nlm, nq = int(h[0]), int(h[1]), int(h[2])
        sp.append(dict(na=na, npr=nlm, lproj=_load(prefix, f"us{i}.l.i32", np.int32),
                       twnl=_load(prefix, f"us{i}.betag.f64", np.float64).reshape(nlm, b["ngw"]),
                       tau=_load(prefix, f"us{i}.tau.f64", np.float64).reshape(na, 3),
                       lm1=_load(prefix, f"us{i}.lm1.i32", np.int32), lm2=_load(prefix, f"us{i}.lm2.i32", np.int32),
                       qaug=_load(prefix, f"us{i}.qaug.f64", np.float64),
                       betapsi=_load(prefix, f"us{i}.betapsi.f64", np.complex128).reshape(case.nst, na * nlm)))
    r["species"] = sp
    r["spsi"] = _load(prefix, "spsi.f64", np.complex128).reshape(case.nst, b["mloc"])
    return r


def run_reference_usx(case: Case, seed: int = 1, occ=None, workdir: str | None = None, threads: int = 8) -> dict:
    """SURVEY section 8 row f4, remainder: ``ref_driver usx`` = the reference's own ultrasoft branch of
    NonLocalPotential::energy (NonLocalPotential.cc:1554-1752) and the augmentation charges of
    ChargeDensity::update_density (ChargeDensity.cc:312-465) on a Sample set up as the application does, with seeded
    coefficients, the given occupations and a seeded effective potential veff(G) on the density basis.  Returns the
    wavefunction basis, the density basis (forced complex at k = 0 for ultrasoft potentials, ChargeDensity.cc:75), per species
    the betag tables / channels / (lm1, lm2, dzero) / Q_nm(G) on the density basis / positions / betapsi, and the results:
    rho_nc (no augmentation), rho (with), nelectrons, veff, enl, hnl."""
    assert have_ref(), "oracle/_ref/ref_driver missing: run `make -C oracle ref` where /root/reference exists"
    tmp = workdir or tempfile.mkdtemp(prefix="qbrefusx_")
    prefix = os.path.join(tmp, "case")
    cf = os.path.join(tmp, "case.txt")
    with open(cf, "w") as f:
        f.write(case.text(prefix))
    env = dict(os.environ, OMP_NUM_THREADS=str(threads))
    subprocess.run([REF_DRIVER, "basis", cf], check=True, env=env, stdout=subprocess.DEVNULL)
    b = read_basis(prefix)
    c = synth_coefficients(b["kpg2"], case.ecut, case.nst, b["mloc"], b["is_real"], seed)
    occ = np.full(case.nst, 2.0) if occ is None else np.asarray(occ, dtype=np.float64)
    c.tofile(prefix + ".in_c.f64")
    occ.tofile(prefix + ".in_occ.f64")
    subprocess.run([REF_DRIVER, "usx", cf], check=True, env=env, stdout=subprocess.DEVNULL)
    r = dict(b)
    h = _load(prefix, "usx.hdr.i32", np.int32)
    assert int(h[6]) == b["ngw"] and int(h[7]) == b["mloc"]
    ngv, nvr = int(h[3]), int(h[4])
    vr = _load(prefix, "usx.vrods.i32", np.int32).reshape(4, nvr)
    mm = _load(prefix, "usx.vidxmm.i32", np.int32)
    r.update(c=c, occ=occ, vgrid=(int(h[0]), int(h[1]), int(h[2])), ngv=ngv, v_is_real=bool(h[5]),
             v_rod_h=vr[0].copy(), v_rod_k=vr[1].copy(), v_rod_lmin=vr[2].copy(), v_rod_size=vr[3].copy(),
             v_idxmin1=int(mm[0]), v_idxmax1=int(mm[1]),
             vkpgx=_load(prefix, "usx.vkpgx.f64", np.float64).reshape(3, ngv), vg2=_load(prefix, "usx.vg2.f64", np.float64),
             rho_nc=_load(prefix, "usx.rho_nc.f64", np.float64), rho=_load(prefix, "usx.rho.f64", np.float64),
             nelectrons=float(_load(prefix, "usx.nel.f64", np.float64)[0]), veff=_load(prefix, "usx.veff.f64", np.complex128),
             enl=float(_load(prefix, "usx.enl.f64", np.float64)[0]),
             hnl=_load(prefix, "usx.hnl.f64", np.complex128).reshape(case.nst, b["mloc"]))
    sp = []
    for i in range(b["nsp"]):
        hh = _load(prefix, f"usx{i}.hdr.i32", np.int32)
        na, nlm, nq = int(hh[0]), int(hh[1]), int(hh[2])
        sp.append(dict(na=na, npr=nlm, nq=nq, lproj=_load(prefix, f"usx{i}.l.i32", np.int32),
                       twnl=_load(prefix, f"usx{i}.betag.f64", np.float64).reshape(nlm, b["ngw"]),
                       tau=_load(prefix, f"usx{i}.tau.f64", np.float64).reshape(na, 3),
                       lm1=_load(prefix, f"usx{i}.lm1.i32", np.int32), lm2=_load(prefix, f"usx{i}.lm2.i32", np.int32),
                       dzero=_load(prefix, f"usx{i}.dzero.f64", np.float64),
                       qnmg=_load(prefix, f"usx{i}.qnmg.f64", np.complex128).reshape(nq, ngv),
                       betapsi=_load(prefix, f"usx{i}.betapsi.f64", np.complex128).reshape(case.nst, na * nlm),
                       wt=np.zeros(nlm)))
    r["species"] = sp
    return r
