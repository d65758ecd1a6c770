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
def splitmix_uniform(seed: int, start: int, count: int) -> np.ndarray:
    """u[i] in [0,1) for counters start..start+count-1 (splitmix64 finaliser, top 53 bits)."""
    with np.errstate(over="ignore"):
        z = (np.arange(start, start + count, dtype=np.uint64) + np.uint64(seed)) * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return (z >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def synth_coefficients(kpg2: np.ndarray, ecut: float, nst: int, ldc: int, is_real: bool, seed: int = 1,
                       first_state: int = 0) -> np.ndarray:
    """c[n, ig] = (u1-1/2 + i(u2-1/2)) * exp(-|k+G|^2 / (ecut/2)) ; Im c[G=0] = 0 for real bases; padding rows zero.

    Returned as a C-contiguous (nst, ldc) complex128 array == the reference's column-major mloc x nst block.
    State n uses counters based on its GLOBAL index first_state+n, so shards regenerate consistently."""
    ngw = kpg2.shape[0]
    damp = np.exp(-kpg2 / (0.5 * ecut))
    c = np.zeros((nst, ldc), dtype=np.complex128)
    for n in range(nst):
        g = first_state + n
        u = splitmix_uniform(seed, 2 * ngw * g, 2 * ngw)
        c[n, :ngw] = ((u[0::2] - 0.5) + 1j * (u[1::2] - 0.5)) * damp
    if is_real:
        c[:, 0] = c[:, 0].real  # G=0 is ig=0 on rod(0,0) for real bases (SlaterDet.cc:2776-2779)
    return c


def synth_potential(np0: int, np1: int, np2: int, seed: int = 7) -> np.ndarray:
    """smooth periodic v(r) (three cosines, O(1) Ha) + 1e-3 fixed-seed noise; x fastest (FourierTransform.h:165)."""
    x = np.arange(np0) / np0
    y = np.arange(np1) / np1
    z = np.arange(np2) / np2
    v = (-0.7 * np.cos(2 * np.pi * x)[None, None, :] - 0.5 * np.cos(2 * np.pi * (y + 0.1))[None, :, None]
         - 0.3 * np.cos(4 * np.pi * (z + 0.2))[:, None, None])
    v = v + 0.25 * np.cos(2 * np.pi * (x[None, None, :] + y[None, :, None] - z[:, None, None]))
    n = np0 * np1 * np2
    v = v.reshape(n) + 1e-3 * (splitmix_uniform(seed, 0, n) - 0.5)
    return np.ascontiguousarray(v)


def synth_occ(nst: int, nocc: int | None = None) -> np.ndarray:
    nocc = nst if nocc is None else nocc
    occ = np.zeros(nst)
    occ[:nocc] = 2.0
    if nocc < nst:  # a fractional level so fac>0 filtering and weights are both exercised
        occ[nocc - 1] = 1.25
    return occ


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
    sp = []
    for i in range(b["nsp"]):
        h = _load(prefix, f"sp{i}.hdr.i32", np.int32)
        na, npr = int(h[0]), int(h[1])
        sp.append(dict(na=na, npr=npr, lproj=_load(prefix, f"sp{i}.lproj.i32", np.int32),
                       wt=_load(prefix, f"sp{i}.wt.f64", np.float64),
                       twnl=_load(prefix, f"sp{i}.twnl.f64", np.float64).reshape(npr, b["ngw"]),
                       tau=_load(prefix, f"sp{i}.tau.f64", np.float64).reshape(na, 3)))
    r["species"] = sp
    if b["nsp"]:
        r["hnl"] = _load(prefix, "hnl.f64", np.complex128).reshape(case.nst, b["mloc"])
    assert r["bwd0"].shape[0] == N
    return r
