#!/usr/bin/env python
"""bench.py -- H psi + density throughput of the B200 path on the reference's headline shapes (BASELINE.json).

A "step" = one pass of the hot path over the rank's state block: the whole H psi block (nonlocal projectors + kinetic
+ local v(r) round trip, EnergyFunctional::energy compute_hpsi=true) followed by one density build
(ChargeDensity::update_density: compute_density over the block, all-reduce of rho over ranks).
value = state-applies/s = states through (H psi + density) per second, whole job.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload mgo216|au992|sih4] [--impl ours|reference]

N>1: launched by torchrun, one rank per GPU; states are sharded by rank (band parallelism, nprow=1); the per-GPU shard is
fixed (weak scaling); the only collectives are the NCCL all-reduce of rho(r) and of the scalar E_nl.
"""
from __future__ import annotations

import argparse
import json
import os
import shutil
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np
from ctypes import byref as C_byref, c_double as C_double

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC, UNIT = "hpsi_density_state_applies_per_s", "state-applies/s"
_COMM = {}               # the process's qb200 communicator (N > 1), shared by the sub-records
_FP64_PEAK = {}          # device index -> (DMMA TFLOP/s, DFMA TFLOP/s), measured once per process

WORKLOADS = {
    # name: cell, ecut (Ha), kpoint, force_complex, nst per GPU, species shape [(name, na, lproj)], note
    "mgo216": dict(cell=(23.1, 0, 0, 0, 23.1, 0, 0, 0, 23.1), ecut=25.0, kpoint=(0, 0, 0), force_complex=True, nst=768,
                   species=[("Mg", 108, [0, 1, 1, 1]), ("O", 108, [0])],
                   note="examples/MgO216 (timing_nr512.i: 50 Ry, force_complex_wf ON, 768 states, 112^3 grid)"),
    "au992": dict(cell=(30.84, 0, 0, 0, 30.84, 0, 0, 0, 119.505), ecut=65.0, kpoint=(1e-7, 0, 0), force_complex=False, nst=256,
                  species=[("Au", 992, [0, 1, 1, 1])],
                  note="examples/gold_benchmark N=992-equivalent cell (252x252x896 grid); per-GPU shard of 256 states (the 5456-state job of BASELINE.json holds 682 per GPU on 8, the 10118-state job of the shipped Au_PBE.xml 1265; c + H psi of 256 states = 23 GB)"),
    "si54p": dict(cell=(0, 15.525, 15.525, 15.525, 0, 15.525, 15.525, 15.525, 0), ecut=32.5, kpoint=(0, 0, 0), force_complex=False, nst=109,
                  species=[("Si", 54, [0, 1, 1, 1])],
                  note="examples/si54p as a Gamma-point real-wavefunction case (65 Ry, 126^3 grid, 109 states, 216 projectors; SURVEY.md 8d)"),
    "sih4": dict(cell=(14, 0, 0, 0, 14, 0, 0, 0, 14), ecut=18.0, kpoint=(0, 0, 0), force_complex=False, nst=4,
                 species=[("Si", 1, [0, 1, 1, 1])], note="examples/sih4 (Gamma, real wavefunctions, 60^3 grid)"),
}


# ------------------------------------------------------------------------------------------------ host-side basis
def make_basis(cell, ecut, kpoint, force_complex):
    """G-sphere / rod tables as Basis::resize builds them (host setup, stays on the CPU in the reference too)"""
    from qball_b200 import basis as B
    return B.make_basis(cell, ecut, kpoint, force_complex), B.density_grid(cell, ecut)


def atom_positions(cell, na_total):
    """simple cubic filling of the cell (positions only steer the phases e^{-iG.tau}, not the cost)"""
    a = np.array(cell, dtype=np.float64).reshape(3, 3)
    n = int(np.ceil(na_total ** (1.0 / 3.0)))
    pts = []
    for i in range(n):
        for j in range(n):
            for k in range(n):
                pts.append(((i + 0.25) / n, (j + 0.25) / n, (k + 0.25) / n))
    pts = np.array(pts[:na_total]) @ a
    return pts


def synth_species(b, wl, seed=11):
    """Gaussian-times-polynomial stand-in projector tables of the right shape"""
    rng = np.random.default_rng(seed)
    na_total = sum(s[1] for s in wl["species"])
    pos = atom_positions(wl["cell"], na_total)
    out, o = [], 0
    kpg = np.sqrt(b["kpg2"])
    for name, na, lproj in wl["species"]:
        npr = len(lproj)
        twnl = np.empty((npr, b["ngw"]))
        for i, l in enumerate(lproj):
            ang = 1.0 if l == 0 else b["kpgx"][(i - 1) % 3] / np.maximum(kpg, 1e-12)
            twnl[i] = (1.5 + 0.1 * i) * kpg ** l * np.exp(-b["kpg2"] / 6.0) * ang
        out.append(dict(na=na, npr=npr, lproj=np.array(lproj, dtype=np.int32), wt=rng.uniform(0.5, 2.0, npr) * np.where(np.array(lproj) == 0, 1, -1),
                        twnl=twnl, tau=pos[o:o + na].copy()))
        o += na
    return out


# ------------------------------------------------------------------------------------------------ clocks sampler
class ClockSampler:
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ reference arm / cpu baseline
def reference_case(wl_name, nst_sample, tmp):
    """the workload as a case of oracle/_ref/ref_driver: synthetic species files of the workload's projector shape
    (oracle/synth_species.py; /root/reference and its pseudopotentials do not exist on the GPU box) and the same atom
    positions as the GPU arm"""
    import refdrive as R
    import synth_species as S
    wl = WORKLOADS[wl_name]
    spfiles = {"mgo216": S.mgo_species, "au992": S.au_species, "si54p": S.si_species}.get(wl_name)
    species, atoms = [], []
    if spfiles:
        species = spfiles(tmp)
        na_total = sum(s[1] for s in wl["species"])
        pos = atom_positions(wl["cell"], na_total)
        o = 0
        for (nm, _), (_, na, _) in zip(species, wl["species"]):
            for i in range(na):
                atoms.append((f"{nm}{i}", nm, float(pos[o + i][0]), float(pos[o + i][1]), float(pos[o + i][2])))
            o += na
    return R.Case(cell=wl["cell"], ecut=wl["ecut"], kpoint=wl["kpoint"], force_complex=wl["force_complex"], nst=nst_sample,
                  species=species, atoms=atoms)


def run_reference_cpu(wl_name, nst_sample, nrep, threads=None):
    """times the reference's own SlaterDet::rs_mul_add / compute_density / NonLocalPotential::energy (oracle/_ref, the
    UNMODIFIED reference compiled serially with its built-in FFT) on a bounded sample of the workload."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import refdrive as R
    wl = WORKLOADS[wl_name]
    threads = threads or os.cpu_count() or 1
    if not R.have_ref():
        return None
    tmp = tempfile.mkdtemp(prefix="qbbench_")
    case = reference_case(wl_name, nst_sample, tmp)
    prefix = os.path.join(tmp, "case")
    cf = os.path.join(tmp, "case.txt")
    with open(cf, "w") as f:
        f.write(case.text(prefix))
    env = dict(os.environ, OMP_NUM_THREADS=str(threads))
    subprocess.run([R.REF_DRIVER, "basis", cf], check=True, env=env, stdout=subprocess.DEVNULL)
    b = R.read_basis(prefix)
    R.synth_coefficients(b["kpg2"], wl["ecut"], nst_sample, b["mloc"], b["is_real"], 1).tofile(prefix + ".in_c.f64")
    R.synth_potential(b["np0"], b["np1"], b["np2"], 7).tofile(prefix + ".in_v.f64")
    R.synth_occ(nst_sample).tofile(prefix + ".in_occ.f64")
    out = subprocess.run([R.REF_DRIVER, "time", cf, str(nrep)], check=True, env=env, capture_output=True, text=True).stdout
    line = [l for l in out.splitlines() if l.startswith("{")][-1]
    t = json.loads(line)
    per_rep = (t["t_nonlocal"] + t["t_kinetic"] + t["t_local"] + t["t_density"]) / nrep
    t.update(per_rep_s=per_rep, applies_per_s=nst_sample / per_rep, threads=threads)
    shutil.rmtree(tmp, ignore_errors=True)
    return t


def parity_vs_reference(wl_name, nst_sample, device, stream):
    """The oracle as the CHECKER of the GPU arm on the benchmark's own configuration: the compiled reference
    (oracle/_ref/ref_driver `run`) evaluates H psi, rho and E_nl for the first `nst_sample` states of the workload -- the
    same seeded coefficients / potential / atom positions the CPU timing leg uses, all atoms' projector tables as the
    reference's own NonLocalPotential::update_twnl builds them from the species files -- and the GPU path (C ABI) is run
    on exactly those inputs and tables.  Returns max-norm relative errors (north_star gate: 1e-10)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import torch
    import refdrive as R
    from qball_b200 import host as H
    if not R.have_ref():
        return None
    wl = WORKLOADS[wl_name]
    tmp = tempfile.mkdtemp(prefix="qbparity_")
    try:
        t0 = time.perf_counter()
        r = R.run_reference(reference_case(wl_name, nst_sample, tmp), seed=1, nocc=None, workdir=tmp,
                            threads=os.cpu_count() or 1)        # v: seed + 6 = 7, as the timing leg
        t_ref = time.perf_counter() - t0
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    b, _ = make_basis(wl["cell"], wl["ecut"], wl["kpoint"], wl["force_complex"])
    assert b["ngw"] == r["ngw"] and np.array_equal(b["rod_h"], r["rod_h"]) and np.array_equal(b["idx"], r["idx"]), "basis tables differ from the reference's"
    ft = H.FourierTransform(b, r["np0"], r["np1"], r["np2"], device=device, stream=stream)
    nlp = H.NonLocalPotential(b, r["species"], device=device, stream=stream)
    N = r["np0"] * r["np1"] * r["np2"]
    dev = torch.device("cuda", device)
    with torch.cuda.stream(stream):
        cd = torch.from_numpy(r["c"]).to(dev)
        vd = torch.from_numpy(r["v"]).to(dev)
        out = torch.zeros_like(cd)
        enl = H.hpsi(ft, nlp, cd, r["occ"], vd, torch.from_numpy(b["kpg2"]).to(dev), out)
        rho = torch.zeros(N, dtype=torch.float64, device=dev)
        H.compute_density(ft, cd, 1.0, r["occ"], r["omega"], rho)
    torch.cuda.synchronize(dev)
    hp, rh = out.cpu().numpy(), rho.cpu().numpy()
    res = {"hpsi_relerr": float(np.abs(hp - r["hpsi"]).max() / np.abs(r["hpsi"]).max()),
           "rho_relerr": float(np.abs(rh - r["rho"]).max() / np.abs(r["rho"]).max()),
           "enl_abs": float(abs(enl - r["enl"])), "enl_rel": float(abs(enl - r["enl"]) / max(1.0, abs(r["enl"]))),
           "enl": float(enl), "enl_reference": float(r["enl"]), "states": nst_sample,
           "projector_rows": int(sum(s["na"] * s["npr"] for s in r["species"])), "projector_form": int(nlp.query(14)),
           "reference_seconds": t_ref,
           "what": f"first {nst_sample} states of the {wl_name} workload, all atoms' projectors: qb200_hpsi + qb200_compute_density "
                   "against oracle/_ref/ref_driver (unmodified reference: NonLocalPotential::energy + kinetic + SlaterDet::rs_mul_add, "
                   "SlaterDet::compute_density) on identical inputs and projector tables; tolerance 1e-10"}
    res["ok"] = bool(res["hpsi_relerr"] < 1e-10 and res["rho_relerr"] < 1e-10 and res["enl_rel"] < 1e-10)
    del ft, nlp, cd, vd, out, rho
    torch.cuda.empty_cache()
    return res


def probe_cpu_fft_libraries():
    """BASELINE.md tier B (the reference built against FFTW3) needs an FFTW3 on the box: probed, not assumed"""
    import ctypes.util
    found = {name: bool(ctypes.util.find_library(name)) for name in ("fftw3", "fftw3_omp", "fftw3_threads", "mkl_rt", "essl")}
    found["note"] = "no FFTW3/MKL/ESSL on this image: tier B (reference + FFTW3) cannot be built; the CPU arm is the reference's built-in FFT" \
        if not any(v for v in found.values() if isinstance(v, bool)) else "an FFT library is present: rebuild oracle/_ref with HAVE_FFTW3 for the tier-B arm"
    return found


def norm2_states(c, ngw, is_real):
    """sum_G |c_G|^2 per state over the FULL sphere (real bases store half of it: G and -G, G = 0 once)"""
    a = (c[:, :ngw].real ** 2 + c[:, :ngw].imag ** 2).sum(axis=1)
    return 2.0 * a - c[:, 0].real ** 2 if is_real else a


def main():
    # the contract is ONE JSON line on stdout: keep a private handle to the real stdout and point fd 1 at stderr, so that
    # library chatter (NCCL prints its version banner to stdout) cannot get in front of it
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)

    def emit(obj):
        real_stdout.write(json.dumps(obj) + "\n")
        real_stdout.flush()

    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="mgo216", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--nst", type=int, default=0, help="states per GPU (default: the workload's)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workspace-mb", type=int, default=0, help="qb200_plan_set_workspace (0: the library default)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: the workload's states PER GPU (default); strong: the workload's states IN TOTAL, split over the GPUs")
    ap.add_argument("--no-sub", action="store_true", help="skip the au992 / strong-scaling sub-records of the default run")
    args = ap.parse_args()
    wl = dict(WORKLOADS[args.workload])
    if args.nst:
        wl["nst"] = args.nst
    if args.scaling == "strong":
        wl["nst"] = max(1, wl["nst"] // int(os.environ.get("WORLD_SIZE", "1")))
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    metric, unit = METRIC, UNIT
    config = {"workload": f"{args.workload}: {wl['note']}", "states_per_gpu": wl["nst"], "sharding": f"band x{world}" + ("; strong-scaling run: one NCCL all-reduce per step carries rho(r) and E_nl, no host synchronisation inside the step" if world > 1 and args.scaling == "strong" else ""),
              "l2": "inputs larger than L2 (coefficient block >> 126 MB)" if args.workload == "mgo216" else "L2 flushed by the c/Hpsi block sweep",
              "anl": "the materialised projector block anl(G) depends on the atomic positions only and is kept across the steps (positions fixed, as over the SCF iterations of one ionic step); QB200_ANL_CACHE=0 regenerates it in every call like the reference's comp_anl (+0.08 ms per MgO216 step)"}

    # ---------------------------------------------------------------- reference arm: the reference's CPU path on host cores
    if args.impl == "reference":
        if rank != 0:
            return
        sample = {"mgo216": 16, "au992": 1, "sih4": 4, "si54p": 16}[args.workload]
        ts = []
        for i in range(args.warmup + args.steps):
            t = run_reference_cpu(args.workload, sample, 1)
            if t is None:
                emit({"impl": "reference", "unavailable": "oracle/_ref/ref_driver not built"})
                return
            if i >= args.warmup:
                ts.append(t)
            if i == 0 and t["per_rep_s"] > 60:   # keep the whole run within minutes
                args.warmup, args.steps = 0, max(1, min(args.steps, 2))
                ts.append(t)
            if len(ts) >= args.steps:
                break
        per = float(np.mean([t["per_rep_s"] for t in ts]))
        val = sample / per
        smp = f"{sample} states of the {args.workload} shape (all {sum(s[1] for s in wl['species'])} atoms' projectors), one pass = " \
              f"NonLocalPotential::energy + kinetic + rs_mul_add + compute_density, built-in FFT (FFT_NOLIB)"
        emit(({"impl": "reference", "metric": metric, "value": val, "unit": unit, "n_gpus": args.gpus, "steps": len(ts),
                          "warmup": args.warmup, "ms_per_step": per * 1e3, "higher_is_better": True, "scaling": "weak",
                          "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
                          "cpu_baseline": {"value": val, "unit": unit, "cores": ts[0]["threads"], "kind": "reference", "sample": smp},
                          "e2e": {"value": val, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                          "breakdown_s": {k: float(np.mean([t[k] for t in ts])) for k in ("t_nonlocal", "t_kinetic", "t_local", "t_density")}}))
        return

    # ---------------------------------------------------------------- our arm
    import torch
    import torch.distributed as dist
    from qball_b200 import capi
    from qball_b200 import host as H
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    # run (and first-touch the pinned host blocks) on the cores NVML reports as closest to this rank's GPU, so that the
    # host<->device traffic of the e2e path stays on the local NUMA node / PCIe root when several ranks share the box
    try:
        import pynvml
        pynvml.nvmlInit()
        hnd = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(hnd, (ncpu + 63) // 64)
        cpus = [w * 64 + bit for w, word in enumerate(words) for bit in range(64) if (word >> bit) & 1 and w * 64 + bit < ncpu]
        if cpus and world > 1:
            os.sched_setaffinity(0, cpus)
    except Exception as ex:  # noqa: BLE001
        sys.stderr.write(f"cpu affinity not set: {ex}\n")
    if world > 1:
        import datetime
        # every collective of this bench is short: a rank that dies must not hold the others for NCCL's default 10 minutes
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=150))
    out = run_ours(args, args.workload, wl, rank, world, local_rank, args.steps, args.warmup, extras=not args.no_e2e,
                   scaling=args.scaling)
    # ---------------------------------------------------------------- sub-records of the default run (the driver runs
    # `bench.py --gpus N --steps K --warmup W` only): the other half of the metric (Au992 shape, fixed 64-state shard per
    # GPU) and, at N > 1, MgO216 STRONG scaling (768 states in total, 768/N per GPU) next to the weak-scaling headline
    if args.workload == "mgo216" and not args.nst and not args.no_sub and args.scaling == "weak":
        keep = ("value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "scaling", "config", "gpu_launches", "roofline",
                "roofline_hbm", "roofline_local_path", "roofline_fp64", "kernel_ms_per_step", "shape", "parity", "enl", "e2e", "scf_iteration")
        if world > 1:
            wls = dict(WORKLOADS["mgo216"])
            wls["nst"] = WORKLOADS["mgo216"]["nst"] // world
            try:
                sub = run_ours(args, "mgo216", wls, rank, world, local_rank, args.steps, args.warmup, extras=False, want_e2e=True,
                               scaling="strong")
                out["strong_scaling"] = {k: sub[k] for k in keep if k in sub}
            except Exception as ex:  # noqa: BLE001
                out["strong_scaling"] = {"error": str(ex)[:300]}
        try:
            sub = run_ours(args, "au992", dict(WORKLOADS["au992"]), rank, world, local_rank, max(3, min(args.steps, 4)), 1, extras=False)
            out["au992"] = {k: sub[k] for k in keep if k in sub}
        except Exception as ex:  # noqa: BLE001
            out["au992"] = {"error": str(ex)[:300]}
    if rank == 0:
        emit(out)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


def run_ours(args, wl_name, wl, rank, world, local_rank, steps, warmup, extras, want_e2e=False, scaling="weak"):
    """one measurement of the B200 arm on workload `wl` (per-GPU shard wl["nst"]); extras: the side measurements of the
    headline record (TDDFT glue, cuFFT comparison, subspace LA, e2e, CPU baseline, parity against the reference)"""
    import torch
    import torch.distributed as dist
    from qball_b200 import capi
    from qball_b200 import host as H
    dev = torch.device("cuda", local_rank)
    metric, unit = METRIC, UNIT
    config = {"workload": f"{wl_name}: {wl['note']}", "states_per_gpu": wl["nst"], "sharding": f"band x{world}" + ("; strong-scaling run: one NCCL all-reduce per step carries rho(r) and E_nl, no host synchronisation inside the step" if world > 1 and scaling == "strong" else ""),
              "l2": "inputs larger than L2 (coefficient block >> 126 MB)" if wl["nst"] * 16 * 70000 > (200 << 20) else "L2 flushed by the c/Hpsi block sweep",
              "anl": "the materialised projector block anl(G) depends on the atomic positions only and is kept across the steps (positions fixed, as over the SCF iterations of one ionic step); QB200_ANL_CACHE=0 regenerates it in every call like the reference's comp_anl (+0.08 ms per MgO216 step)"}
    b, grid = make_basis(wl["cell"], wl["ecut"], wl["kpoint"], wl["force_complex"])
    np0, np1, np2 = grid
    N, ngw, nst = np0 * np1 * np2, b["ngw"], wl["nst"]
    from qball_b200 import synth as R
    # the block is synthesised directly in HBM (bit-identical to the numpy generator the CPU arm and the fixtures use);
    # a host copy exists only where a host-pointer leg needs one
    need_host_block = extras or want_e2e
    c_dev0 = R.synth_coefficients_torch(b["kpg2"], wl["ecut"], nst, ngw, b["is_real"], seed=1, first_state=rank * nst, device=dev)
    c_host = c_dev0.cpu().numpy() if need_host_block else None
    v_host = R.synth_potential(np0, np1, np2, 7)
    occ = R.synth_occ(nst, nst - max(1, nst // 64))
    species = synth_species(b, wl)
    stream = torch.cuda.Stream(device=dev)
    ft = H.FourierTransform(b, np0, np1, np2, device=local_rank, stream=stream)
    nlp = H.NonLocalPotential(b, species, device=local_rank, stream=stream)
    if args.workspace_mb:
        ft.set_workspace(args.workspace_mb << 20)
    with torch.cuda.stream(stream):
        c = c_dev0
        v = torch.from_numpy(v_host).to(dev)
        kpg2 = torch.from_numpy(b["kpg2"]).to(dev)
        hpsi = torch.zeros_like(c)
        rho_ext = torch.zeros(N + 8, dtype=torch.float64, device=dev)     # rho(r) and, behind it, the scalars that ride in the same all-reduce
        rho = rho_ext[:N]
        scal = torch.zeros(2, dtype=torch.float64, device=dev)
    omega = b["omega"]

    # N > 1: the collectives of the C ABI itself (NCCL inside libqball_b200.so; torch.distributed is only the rendezvous)
    comm = _COMM.get("comm")
    if world > 1 and comm is None:
        from qball_b200 import parallel as PAR
        comm = _COMM["comm"] = PAR.Communicator.from_torch_distributed(local_rank)

    # N > 1, two ways to run the step's collectives (A/B on one 8-GPU box, tools/gpu_n_ab.sh, profiles/r2z3_allreduce_ab_8gpu.txt):
    #  * E_nl returned to the host by qb200_hpsi, then qb200_allreduce_rho + qb200_allreduce_scalars (two host synchronisations per step);
    #  * fused: no host synchronisation, E_nl placed behind rho(r) on the device (qb200_nl_last_enl) and ONE all-reduce.
    # Fused wins where the step is short (96 states per GPU: 2.83 against 3.01 ms at 8 GPUs; 2 GPUs, 768 states: 19.32 against
    # 19.51-19.62 ms) and loses at 8 GPUs with long steps (768 states per GPU: 20.6-21.1 against 19.59 ms; cause not identified
    # within the round's GPU budget), so it is the default of the strong-scaling run only; QB200_BENCH_FUSED_ALLREDUCE=0/1 overrides.
    fused_allreduce = os.environ.get("QB200_BENCH_FUSED_ALLREDUCE", "1" if scaling == "strong" else "0") == "1"

    def step():
        with torch.cuda.stream(stream):
            if world == 1:
                enl = H.hpsi(ft, nlp, c, occ, v, kpg2, hpsi)
                rho.zero_()
                H.compute_density(ft, c, 1.0, occ, omega, rho)
            elif not fused_allreduce:
                enl = H.hpsi(ft, nlp, c, occ, v, kpg2, hpsi)
                rho.zero_()
                H.compute_density(ft, c, 1.0, occ, omega, rho)
                comm.allreduce_rho(rho, stream)
                enl = comm.allreduce_scalars([enl])[0]
            else:
                # no host synchronisation inside the step: E_nl stays on the device, is placed behind rho(r), and ONE all-reduce
                # carries both (ChargeDensity.cc:309 dsum('r') of rho; NonLocalPotential.cc:2629 / EnergyFunctional.cc:1519 for E_nl)
                H.hpsi(ft, nlp, c, occ, v, kpg2, hpsi, want_enl=False)
                rho_ext.zero_()
                H.compute_density(ft, c, 1.0, occ, omega, rho)
                nlp.last_enl(rho_ext[N:N + 1])
                comm.allreduce_rho(rho_ext, stream)
                enl = None
        return enl

    def sync_all():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    for _ in range(warmup):
        step()
    sync_all()
    l0 = ft.launches() + nlp.launches()
    capi.profile_read()
    capi.profile_enable(True)
    sampler = ClockSampler(local_rank)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e0.record(stream)
    for _ in range(steps):
        enl = step()
    with torch.cuda.stream(stream):
        e1.record(stream)
    sync_all()
    if world > 1 and enl is None:
        enl = float(rho_ext[N].item())          # the all-reduced E_nl of the last step, read after the timed region
    clocks = sampler.stop()
    ms = e0.elapsed_time(e1)
    capi.profile_enable(False)
    prof = capi.profile_read()
    launches = ft.launches() + nlp.launches() - l0
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    ms_per_step = ms / steps
    value = world * nst * steps / (ms * 1e-3)

    # ---------------------------------------------------------------- integrity of the (all-reduced) density of the last step:
    # integral of rho = sum_n occ_n ||c_n||^2 (ChargeDensity.cc:525), and at N > 1 the all-reduced rho against the
    # rank-ordered sum of the per-rank densities gathered on every rank (NCCL's reduction order may differ: ~1e-16)
    with torch.cuda.stream(stream):
        nel = float(rho.sum().item()) * omega / N
        nrm = (c[:, :ngw].real ** 2 + c[:, :ngw].imag ** 2).sum(dim=1)
        if b["is_real"]:
            nrm = 2.0 * nrm - c[:, 0].real ** 2           # half sphere stored: G and -G, G = 0 once
        want = (torch.from_numpy(np.asarray(occ, dtype=np.float64)).to(dev) * nrm).sum().reshape(1)
        rho_sum_err = None
        if world > 1:
            dist.all_reduce(want)
            mine = torch.zeros(N, dtype=torch.float64, device=dev)
            H.compute_density(ft, c, 1.0, occ, omega, mine)
            parts = torch.empty((world, N), dtype=torch.float64, device=dev)
            dist.all_gather_into_tensor(parts, mine)
            tot = parts[0].clone()
            for r in range(1, world):
                tot += parts[r]
            rho_sum_err = float(((rho - tot).abs().max() / tot.abs().max()).item())
            del mine, parts, tot
    integrity = {"nelectrons": nel, "sum_occ_norm2": float(want.item()), "nelectrons_relerr": abs(nel - float(want.item())) / float(want.item()),
                 "allreduced_rho_vs_sum_of_rank_rhos_relerr": rho_sum_err}
    integrity["ok"] = bool(integrity["nelectrons_relerr"] < 1e-10 and (rho_sum_err is None or rho_sum_err < 1e-12))

    # ---------------------------------------------------------------- roofline of the dominant kernels (live event times)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak, peak_src = (peaks.get("hbm_gbs"), "measured (MEASURED_PEAKS.json)") if peaks.get("hbm_gbs") else (6650.0, "fallback (B200_PROFILING.md)")
    nvec = ft.nvec()
    nunits_h = (nst // 2 + nst % 2) if b["is_real"] else nst          # FFT units per H psi sweep
    xy_ms, xy_n = prof["xy_stage"]
    # algorithmic bytes of the xy stage per FFT unit: read+write the column-form plane rows, read v (H psi) / rmw rho-partial
    # is charged once per build: H psi unit 32*nvec*np2 + 8*N ; density unit 16*nvec*np2
    xy_bytes_step = nunits_h * (32.0 * nvec * np2 + 8.0 * N) + nst * (16.0 * nvec * np2) + 16.0 * N
    # DRAM traffic per launch from the committed ncu --set full capture of this same command (tools/ncu_traffic.py)
    ncu_traffic = {}
    try:
        ncu_traffic = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
    except Exception:
        pass

    def traffic_of(*keys):
        """launch-weighted mean DRAM bytes per launch over the kernels `keys` (None unless every one was captured)"""
        if wl_name != "mgo216" or not all(k in ncu_traffic for k in keys):
            return None
        return sum(ncu_traffic[k]["dram_bytes_per_launch"] for k in keys) / len(keys)

    roofline_hbm = None
    # the plane kernel that ran: k_plane_t (y direction in tensor memory; plan query 17) or k_plane_s
    plane_keys = ("k_plane_t<0>", "k_plane_t<1>") if ft.query(17) == 1 else ("k_plane_s<0>", "k_plane_s<1>")
    if xy_n:
        ach = xy_bytes_step * steps / (xy_ms * 1e-3) / 1e9
        roofline_hbm = {"kernel": "k_plane (fused xy stage)" if ft.fused() else "k_xrows+k_ycols (split xy stage)", "bound": "hbm",
                        "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak,
                        "traffic": traffic_of(*plane_keys) if ft.fused() else None, "peak_source": peak_src,
                        "algorithmic_bytes_per_launch": xy_bytes_step * steps / xy_n,
                        "launches": xy_n, "avg_launch_ms": xy_ms / xy_n}
        # the stage is bound ON CHIP before it is bound by HBM: the other two roofs, from the committed ncu
        # --set full capture of this command (profiles/ncu_traffic.json; static evidence, not measured in this run)
        if ft.fused() and wl_name == "mgo216":
            onchip = {k: {"fp64_pipe_active_frac": ncu_traffic[k].get("fp64_pipe_active_pct", 0.0) / 100.0,
                          "smem_pipe_wavefronts_per_cycle": ncu_traffic[k].get("smem_wavefronts_per_cycle_per_sm"),
                          "dram_bytes_per_launch": ncu_traffic[k]["dram_bytes_per_launch"]}
                      for k in plane_keys if k in ncu_traffic and "fp64_pipe_active_pct" in ncu_traffic[k]}
            if onchip:
                roofline_hbm["on_chip_roofs"] = {"kernels": onchip, "source": "profiles/ncu_traffic.json (ncu --set full of this command)",
                                                 "note": "the stage is bound by FP64 issue: ~6 ms per step of butterflies at the measured DFMA rate "
                                                         "(k_plane_t keeps the y direction in tensor memory, so the shared-memory pipe is no longer a roof; DESIGN.md section 4)"}
    # the whole local path (z columns + xy stage, both directions) against the HBM roofline with SURVEY.md section 8d's
    # per-unit algorithmic bytes: B_Hpsi = 48*ngw*cper + 64*nvec*np2 + 8*N, B_rho = 16*ngw + 32*nvec*np2 (+16*N per build)
    cper = 2 if b["is_real"] else 1
    local_bytes_step = nunits_h * (48.0 * ngw * cper + 64.0 * nvec * np2 + 8.0 * N) + nst * (16.0 * ngw + 32.0 * nvec * np2) + 16.0 * N
    local_ms = prof["k_zcol_bwd"][0] + prof["xy_stage"][0] + prof["k_zcol_fwd"][0] + prof["k_rho_reduce"][0]
    roofline_local = None
    if local_ms > 0:
        ach = local_bytes_step * steps / (local_ms * 1e-3) / 1e9
        roofline_local = {"kernel": "local path: k_zcol_bwd + xy stage + k_zcol_fwd (H psi local term + density)", "bound": "hbm",
                          "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak, "traffic": None, "peak_source": peak_src,
                          "note": "xy stage is bound by FP64 issue (ncu: FP64 pipe 64-66 % active, DRAM 11 % of peak), the z-column kernels by HBM (5.3-5.8 TB/s)"}
    if roofline_local is not None:
        roofline_local["fp64_roof_note"] = "the butterflies of the local path execute ~0.18 TFLOP per MgO216 step; at the measured DFMA rate that is the second roof of this stage (see roofline_fp64.fp64_tflops_measured)"
    nl_flops_step = 0.0     # flops EXECUTED on the FP64 tensor pipe
    nl_zgemm_flops_step = 0.0   # the same contraction counted as the reference's zgemm/dgemm (8 / 2 flops per MAC)
    nl_mode = nlp.query(14)   # 0 real basis, 1 four-product, 2 three-product (Karatsuba), 3 Gamma-point half sphere (real-function split)
    m3 = nl_mode == 2
    for s in species:   # 2 GEMMs; real MACs per complex MAC: 4 plain embedding, 3 Karatsuba (nonlocal_3m.cuh), 2 at Gamma (two real functions)
        per_mac = {0: 8.0, 1: 16.0, 2: 12.0, 3: 8.0}[nl_mode]
        nl_flops_step += per_mac * s["na"] * s["npr"] * ngw * nst
        nl_zgemm_flops_step += (8.0 if b["is_real"] else 16.0) * s["na"] * s["npr"] * ngw * nst
    nl_ms = prof["k_fnl"][0] + prof["k_back"][0]
    # FP64 ceiling: measured in this process by the library's DMMA / DFMA issue-rate loops (MEASURED_PEAKS.json holds no
    # FP64 figure); the nominal B200 number is kept beside it
    if local_rank not in _FP64_PEAK:
        try:
            _FP64_PEAK[local_rank] = capi.measure_fp64_peak(local_rank)
        except Exception as ex:  # noqa: BLE001
            sys.stderr.write(f"fp64 peak measurement failed: {ex}\n")
            _FP64_PEAK[local_rank] = (None, None)
    dmma_tf, dfma_tf = _FP64_PEAK[local_rank]
    fp64_peak = dmma_tf if dmma_tf else 37.0
    # FP64 issue roof of the xy stage, live: the kernels' FP64 warp-instruction counts are a property of the code (per unit, from the
    # committed ncu capture of this command); the time and the DFMA issue rate (2 flops per lane-instruction) are measured in this run
    if roofline_hbm is not None and ft.fused() and wl_name == "mgo216" and dfma_tf:
        try:
            inst = sum(ncu_traffic[k]["fp64_warp_inst_per_launch"] / ncu_traffic[k]["units_per_launch"] * n_
                       for k, n_ in ((plane_keys[0], nunits_h), (plane_keys[1], nst)))
            lane_rate = inst * 32.0 * steps / (xy_ms * 1e-3)
            peak_rate = dfma_tf * 1e12 / 2.0
            roofline_hbm["fp64_issue"] = {"bound": "fp64 instruction issue", "achieved": lane_rate / 1e12, "peak": peak_rate / 1e12,
                                          "unit": "T lane-instructions/s", "frac": lane_rate / peak_rate, "fp64_warp_instructions_per_step": inst,
                                          "source": "instruction counts per unit from profiles/ncu_traffic.json (ncu of this command); time and DFMA rate measured in this run"}
        except Exception:  # noqa: BLE001
            pass
    roofline_fp64 = None
    if nl_ms > 0:
        ach = nl_flops_step * steps / (nl_ms * 1e-3) / 1e12
        roofline_fp64 = {"kernel": ("k_fnl3 + k_back3 (DMMA projector GEMMs, 3-product complex form: 12 flops per complex MAC)" if m3
                                    else "k_split_pm + k_fnl<1> + k_back<2> (DMMA projector GEMMs over the half sphere: complex states "
                                         "at Gamma projected as two real functions, 8 flops per complex MAC; the split pass and the scatter epilogue included in the time)"
                                    if nl_mode == 3 else "k_fnl + k_back (DMMA projector GEMMs)"), "bound": "tensor", "achieved": ach, "peak": fp64_peak,
                         "unit": "TFLOP/s", "frac": ach / fp64_peak, "traffic": traffic_of("k_fnl3<4>", "k_back3<4>") if m3 else (traffic_of("k_split_pm", "k_fnl<1>", "k_back<2>") if nl_mode == 3 else None),
                         "zgemm_equivalent_tflops": nl_zgemm_flops_step * steps / (nl_ms * 1e-3) / 1e12,
                         "peak_source": ("measured in this run: qb200_measure_fp64_peak (DMMA m8n8k4 issue-rate loop on every SM)" if dmma_tf
                                         else "nominal FP64 37 TFLOP/s (measurement failed)"),
                         "fp64_tflops_measured": {"dmma": dmma_tf, "dfma": dfma_tf}, "fp64_tflops_nominal": 37.0,
                         "launches": prof["k_fnl"][1] + prof["k_back"][1]}
    prof_ms = {k: round(vv[0] / steps, 4) for k, vv in prof.items()}
    dominant = max(prof.items(), key=lambda kv: kv[1][0])[0]
    roofline = roofline_fp64 if dominant in ("k_fnl", "k_back") and roofline_fp64 else roofline_hbm

    # ---------------------------------------------------------------- TDDFT propagation (configs[3]): one 4th-order Taylor
    # exponential exp(-i dt H) on the resident block = 4 H psi applications + axpy chain (ExponentialWavefunctionStepper.cc:51-149)
    tddft = None
    if not b["is_real"] and extras:
        with torch.cuda.stream(stream):
            cprop = c.clone()
            H.exponential(ft, nlp, cprop, occ, v, kpg2, 0.02)        # warm-up (allocates the two work blocks)
            t0e, t1e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0e.record(stream)
            nrep = 2
            for _ in range(nrep):
                H.exponential(ft, nlp, cprop, occ, v, kpg2, 0.02)
            t1e.record(stream)
        sync_all()
        ms_exp = t0e.elapsed_time(t1e) / nrep
        # current density of the block: 3 directions x 2 fused density passes (qb200_compute_current)
        with torch.cuda.stream(stream):
            kpgx_d = torch.from_numpy(np.ascontiguousarray(b["kpgx"])).to(dev)
            cur = torch.zeros((3, N), dtype=torch.float64, device=dev)
            H.compute_current(ft, c, 1.0, occ, omega, kpgx_d, cur)
            t0c, t1c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0c.record(stream)
            H.compute_current(ft, c, 1.0, occ, omega, kpgx_d, cur)
            t1c.record(stream)
        sync_all()
        ms_cur = t0c.elapsed_time(t1c)
        del cur, kpgx_d
        tddft = {"exponential_order4_ms": ms_exp, "current_density_ms": ms_cur, "hpsi_state_applies_per_s": world * nst * 4 / (ms_exp * 1e-3),
                 "what": "qb200_exponential, order 4, Hamiltonian frozen, block resident in HBM (per rank)"}
        del cprop

    # ---------------------------------------------------------------- cuFFT as a timing comparison (north_star): the same
    # local operator cp += FT[v FT^-1 c] per state on dense grids through torch.fft (cuFFT Z2Z 3-D, batched), coefficients
    # already scattered to the grid and no gather timed -- i.e. cuFFT is given LESS work than the fused path does
    cufft = None
    if rank == 0 and extras:
        try:
            nb_c = max(1, min(nst, int((2 << 30) // (16 * N))))          # states per cuFFT batch (<= 2 GiB per buffer)
            with torch.cuda.stream(stream):
                xg = torch.zeros((nb_c, np2, np1, np0), dtype=torch.complex128, device=dev)
                xg.view(torch.float64).normal_()
                v3 = v.view(1, np2, np1, np0)

                def cufft_step():
                    y = torch.fft.ifftn(xg, dim=(1, 2, 3), norm="forward")     # backward: unscaled synthesis
                    y.mul_(v3)
                    return torch.fft.fftn(y, dim=(1, 2, 3), norm="forward")     # forward: 1/N analysis

                cufft_step(); cufft_step()
                ta, tb, tc_, td = (torch.cuda.Event(enable_timing=True) for _ in range(4))
                ta.record(stream)
                for _ in range(3):
                    cufft_step()
                tb.record(stream)
                cpl = torch.zeros_like(c)
                H.rs_mul_add(ft, c, v, cpl, kpg2)
                tc_.record(stream)
                for _ in range(3):
                    H.rs_mul_add(ft, c, v, cpl, kpg2)
                td.record(stream)
            torch.cuda.synchronize(dev)          # rank 0 only: no collective in this section
            units = nst / 2 if b["is_real"] else nst                    # cuFFT would also pack two real states per FFT
            us_cufft = ta.elapsed_time(tb) / 3 / nb_c * 1e3
            us_ours = tc_.elapsed_time(td) / 3 / units * 1e3
            cufft = {"cufft_us_per_transform_pair": us_cufft, "ours_us_per_unit": us_ours, "speedup": us_cufft / us_ours,
                     "what": f"torch.fft (cuFFT) ifftn + v multiply + fftn on {nb_c} dense {np0}x{np1}x{np2} complex128 grids per batch "
                             "(no sphere scatter/gather, no kinetic term) against qb200_rs_mul_add per FFT unit (scatter, pruned "
                             "transforms, v multiply, gather and kinetic term fused)"}
            del xg, cpl
        except Exception as ex:  # noqa: BLE001
            cufft = {"error": str(ex)[:200]}

    # ---------------------------------------------------------------- subspace dense LA on the resident block:
    # the PSD/PSDA descent direction a = c^H Hc, Hc -= c a (with band sharding: after an NCCL all-gather of the state
    # blocks -- the path's one exchange step) and SlaterDet::gram on this rank's block
    subspace = None
    if extras:
        try:
            la = H.SubspaceLA(b, device=local_rank, stream=stream)
            from qball_b200 import parallel as PAR

            def la_step(do_gram):
                with torch.cuda.stream(stream):
                    call = PAR.allgather_states(c, world * nst) if world > 1 else c
                    la.residual(call, hres)
                    if do_gram:
                        la.gram(cg)

            with torch.cuda.stream(stream):
                hres = hpsi.clone()
                cg = c.clone()
            la_step(True)
            sync_all()
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            with torch.cuda.stream(stream):
                ev[0].record(stream)
            la_step(False)
            with torch.cuda.stream(stream):
                ev[1].record(stream)
                cg.copy_(c)
                la.gram(cg)
                ev[2].record(stream)
            sync_all()
            ms_res, ms_gram = ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2])
            cx = 1.0 if b["is_real"] else 4.0            # real flops per (real | complex) multiply-add pair / 2
            fl_res = 2 * 2.0 * cx * (world * nst) * nst * ngw * (2 if b["is_real"] else 1)
            fl_gram = 2 * 2.0 * cx * nst * nst * ngw * (2 if b["is_real"] else 1)
            subspace = {"residual_ms": ms_res, "residual_tflops_gemm_equivalent": fl_res / (ms_res * 1e-3) / 1e12,
                        "gram_ms": ms_gram, "gram_tflops_gemm_equivalent": fl_gram / (ms_gram * 1e-3) / 1e12,
                        "nall": world * nst, "nst": nst,
                        "what": "qb200_residual (incl. the all-gather of c when sharded) and qb200_gram (incl. the copy of c) on the "
                                "resident block; flops counted as the two zgemm/dgemm calls of the reference (full S for gram)"}
            del hres, cg, la
        except Exception as ex:  # noqa: BLE001
            subspace = {"error": str(ex)[:200]}

    # ---------------------------------------------------------------- e2e: through the public API with HOST buffers, copies inside
    # Two ways a host program can own the data, both timed end to end on the wall clock:
    #  (a) e2e_host_blocks: the reference's ownership unchanged -- the coefficient block and H psi live in host memory
    #      (ComplexMatrix::val) and cross PCIe every step (pipelined uploads / downloads, content tag);
    #  (b) e2e: the wavefunction block stays in HBM across the electronic steps, which the device-side stepper makes possible
    #      (qb200_residual + qb200_psda_update + qb200_gram: c never has to visit the host between two H psi evaluations; the
    #      scf_iteration record below runs exactly that loop); per step the host sends v(r) (the Hartree/XC potential it
    #      builds from rho) and receives rho(r), E_nl and the kinetic sums -- pinned host buffers, copies inside the timing.
    e2e = e2e_host = scf_iter = None
    if extras or want_e2e:
        hc = torch.from_numpy(c_host).pin_memory()
        hv = torch.from_numpy(v_host).pin_memory()
        hk = torch.from_numpy(b["kpg2"]).pin_memory()
        hout = torch.empty_like(hc).pin_memory()
        hrho = torch.zeros(N, dtype=torch.float64).pin_memory()

        tag = [0]

        def step_host():
            # one stepper iteration as the SlaterDet/EnergyFunctional shims drive it: the wavefunction
            # changed (new tag) -> H psi uploads it block by block, overlapped with the kernels and the download of
            # finished H psi blocks; the density build of the same wavefunction reuses the device copy
            tag[0] += 1
            ft.set_coefficient_tag(tag[0])
            with torch.cuda.stream(stream):
                e = H.hpsi(ft, nlp, hc, occ, hv, hk, hout)
                hrho.zero_()
                H.compute_density(ft, hc, 1.0, occ, omega, hrho)
            return e

        def step_resident():
            with torch.cuda.stream(stream):
                e = H.hpsi(ft, nlp, c, occ, hv, kpg2, hpsi)                 # v: HOST buffer in (staged by the C ABI)
                ek = H.ekin_sums(ft, c, occ, b["is_real"], kpg2)[0]         # 14 sums come back to the host
                rho.zero_()
                H.compute_density(ft, c, 1.0, occ, omega, rho)
                if world > 1:
                    comm.allreduce_rho(rho, stream)
                    e, ek = comm.allreduce_scalars([e, ek])
                hrho.copy_(rho, non_blocking=True)                          # rho: HOST buffer out
            stream.synchronize()
            return e, ek

        def timed(fn, n):
            fn()
            sync_all()
            t0 = time.perf_counter()
            for _ in range(n):
                fn()
            sync_all()
            tt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            return float(tt.item())

        ne = max(1, min(steps, 3))
        blk = 16 * ngw * nst
        dt = timed(step_host, ne)
        e2e_host = {"value": world * nst * ne / dt, "unit": unit,
                    "h2d_bytes_per_step": int(blk + 8 * N + 8 * ngw + 8 * N), "d2h_bytes_per_step": int(blk + 8 * N + 8),
                    "steps": ne, "api": "qb200_plan_set_coefficient_tag + qb200_hpsi + qb200_compute_density with pinned HOST pointers for c, v, H psi, rho "
                                        "(c uploaded once per step in blocks of states overlapped with compute, H psi blocks downloaded as they finish)"}
        ft.set_coefficient_tag(0)
        nr = max(3, min(steps, 10))
        dt = timed(step_resident, nr)
        e2e = {"value": world * nst * nr / dt, "unit": unit, "h2d_bytes_per_step": int(8 * N + 8 * nst), "d2h_bytes_per_step": int(8 * N + 8 * 15),
               "steps": nr, "wavefunction": "device-resident across steps (the stepper update runs on the device: see scf_iteration)",
               "api": "qb200_hpsi (v from a pinned HOST buffer) + qb200_ekin_sums + qb200_compute_density"
                      + (" + qb200_allreduce_rho / qb200_allreduce_scalars" if world > 1 else "") + "; rho read back into a pinned HOST buffer, E_nl and the kinetic sums returned to the host",
               "host_blocks_variant": e2e_host}
        # ------------------------------------------------------------ one whole electronic (SCF) iteration on the device:
        # BOSampleStepper with wf_dyn PSDA = density + H psi + E_kin, then a = c^H Hc, Hc -= c a (qb200_residual),
        # preconditioner + Anderson update (qb200_psda_update), SlaterDet::gram (qb200_gram); all states on this rank
        if world == 1 or scaling == "strong":
            try:
                from qball_b200 import parallel as PAR
                la2 = H.SubspaceLA(b, device=local_rank, stream=stream)
                with torch.cuda.stream(stream):
                    cw, cl, dl = c.clone(), torch.zeros_like(c), torch.zeros_like(c)
                    if world == 1:
                        la2.gram(cw)
                    else:                     # band-sharded SlaterDet::gram: overlap columns -> sum over ranks -> replicated Cholesky
                        la2.gram_sharded(comm, PAR.allgather_states(cw, world * nst), rank * nst, nst, cw)
                prec = np.where(0.5 * b["kpg2"] < 4.0, 0.5 / 4.0, 0.5 / np.maximum(0.5 * b["kpg2"], 1e-300))   # Preconditioner.cc:47-90, ecutprec 8 Ry
                # the density basis (k = 0, 4 ecut: ChargeDensity.cc:77-81) and Gaussian stand-ins for the ionic tables
                from qball_b200 import basis as BB
                vb = BB.make_basis(wl["cell"], 4.0 * wl["ecut"], (0, 0, 0), False)
                vft = H.FourierTransform(vb, np0, np1, np2, device=local_rank, stream=stream)
                g2 = vb["kpg2"]
                with torch.cuda.stream(stream):
                    g2i_d = torch.from_numpy(np.where(g2 > 0, 1.0 / np.where(g2 > 0, g2, 1.0), 0.0)).to(dev)
                    gx_d = torch.from_numpy(np.ascontiguousarray(vb["kpgx"])).to(dev)
                    vion_d = torch.from_numpy((-(40.0 / omega) * np.exp(-0.35 * g2)).astype(np.complex128)).to(dev)
                    rhops_d = torch.from_numpy((-(float(np.sum(occ)) * world / omega) * np.exp(-0.16 * g2)).astype(np.complex128)).to(dev)
                    rhog_d = torch.zeros(vb["ngw"], dtype=torch.complex128, device=dev)
                    vscf = v.clone()
                it = [0]
                parts = {}

                def scf_step():
                    t = [torch.cuda.Event(enable_timing=True) for _ in range(8)]
                    with torch.cuda.stream(stream):
                        t[0].record(stream)
                        rho.zero_()
                        H.compute_density(ft, cw, 1.0, occ, omega, rho)
                        if world > 1:
                            comm.allreduce_rho(rho, stream)
                        nel = C_double()
                        capi._check(ft._L.qb200_density_finish(vft._h, capi.ptr(rho), omega, capi.ptr(rhog_d), C_byref(nel)), "qb200_density_finish")
                        t[1].record(stream)
                        en = H.update_vhxc(vft, H.XC_LDA, rho, rhog_d, gx_d, g2i_d, vion_d, rhops_d, omega, vscf)     # rho -> v(r), on the device
                        t[2].record(stream)
                        e = H.hpsi(ft, nlp, cw, occ, vscf, kpg2, hpsi)
                        ek = H.ekin_sums(ft, cw, occ, b["is_real"], kpg2)[0]
                        t[3].record(stream)
                        call = PAR.allgather_states(cw, world * nst) if world > 1 else cw
                        la2.residual(call, hpsi)
                        t[4].record(stream)
                        la2.psda_update(cw, hpsi, cl, dl, occ, prec, it[0] > 0, comm if world > 1 else None)
                        t[5].record(stream)
                        if world == 1:
                            la2.gram(cw)
                        else:
                            la2.gram_sharded(comm, PAR.allgather_states(cw, world * nst), rank * nst, nst, cw)
                        t[6].record(stream)
                    stream.synchronize()
                    it[0] += 1
                    for k, (a, bb) in {"density+rhog": (0, 1), "update_vhxc": (1, 2), "hpsi+ekin": (2, 3), "residual": (3, 4),
                                       "psda_update": (4, 5), "gram": (5, 6)}.items():
                        parts[k] = parts.get(k, 0.0) + t[a].elapsed_time(t[bb])
                    return e, ek, en, nel.value

                ns = 3
                scf_step()
                parts.clear()
                dts = timed(scf_step, ns)
                nsteps_timed = ns + 1           # timed() runs one untimed call first; parts accumulate over all of them
                scf_iter = {"ms_per_iteration": dts / ns * 1e3, "state_applies_per_s": world * nst * ns / dts, "iterations": ns,
                            "parts_ms": {k: vv / nsteps_timed for k, vv in parts.items()},
                            "pcie_bytes_per_iteration": 8 * 24,
                            "what": "one electronic iteration of BOSampleStepper / PSDA with NOTHING but scalars crossing PCIe: density + rho(G) "
                                    "(qb200_compute_density, qb200_density_finish) -> v(r) (qb200_update_vhxc, LDA) -> H psi + E_kin -> residual -> "
                                    "preconditioned Anderson update (qb200_psda_update)"
                                    " -> SlaterDet::gram"
                                    + ("" if world == 1 else " (band-sharded: NCCL all-gather of c for the residual and for gram; gram = overlap columns per rank, "
                                       "qb200_allreduce_rho of the overlap, Cholesky replicated on every rank, own columns of c L^-H: qb200_gram_sharded)")}
                # Wavefunction::diag of the block (eigenvalues + rotation), timed once
                if world == 1:
                    with torch.cuda.stream(stream):
                        td0, td1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                        H.hpsi(ft, nlp, cw, occ, vscf, kpg2, hpsi)
                        td0.record(stream)
                        wd, sweeps = la2.diag(cw, hpsi, eigvec=True)
                        td1.record(stream)
                    stream.synchronize()
                    scf_iter["diag_ms"] = td0.elapsed_time(td1)
                    scf_iter["diag_jacobi_sweeps"] = sweeps
                del vft, g2i_d, gx_d, vion_d, rhops_d, rhog_d, vscf
                del la2, cw, cl, dl
            except Exception as ex:  # noqa: BLE001
                scf_iter = {"error": str(ex)[:300]}

    # ---------------------------------------------------------------- cpu baseline (rank 0, N=1): the compiled reference
    cpu_baseline = None
    parity = {"integrity": integrity}
    if rank == 0 and world == 1 and extras and not args.no_cpu_baseline:
        sample = {"mgo216": 16, "au992": 1, "sih4": 4, "si54p": 16}[wl_name]
        try:
            pr = parity_vs_reference(wl_name, sample, local_rank, stream)
            if pr:
                parity.update(pr)
        except Exception as ex:  # noqa: BLE001
            parity["error"] = str(ex)[:300]
            sys.stderr.write(f"parity leg failed: {ex}\n")
        try:
            tcpu = run_reference_cpu(wl_name, sample, 1)
        except Exception as ex:  # noqa: BLE001
            tcpu = None
            sys.stderr.write(f"cpu baseline failed: {ex}\n")
        if tcpu:
            cpu_baseline = {"value": tcpu["applies_per_s"], "unit": unit, "cores": tcpu["threads"], "kind": "reference",
                            "sample": f"{sample} states of the {wl_name} shape, all projectors, one H psi + density pass; "
                                      f"reference compiled serially with its built-in FFT (oracle/_ref), OMP threads = cores",
                            "breakdown_s": {k: tcpu[k] for k in ("t_nonlocal", "t_kinetic", "t_local", "t_density")},
                            "fft_backend": "FFT_NOLIB (the reference's built-in cfftm, its slowest tier)",
                            "tier_b_probe": probe_cpu_fft_libraries()}

    out = {"metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": steps, "warmup": warmup,
           "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "f64",
           "data": "synthetic", "config": config, "clocks": clocks, "gpu_launches": int(launches), "e2e": e2e,
           "roofline": roofline, "roofline_hbm": roofline_hbm, "roofline_local_path": roofline_local, "roofline_fp64": roofline_fp64, "cpu_baseline": cpu_baseline,
           "parity": parity, "e2e_host_blocks": e2e_host, "scf_iteration": scf_iter,
           "kernel_ms_per_step": prof_ms, "tddft": tddft, "subspace_la": subspace, "cufft_comparison": cufft, "enl": enl,
           "shape": {"ngw": ngw, "nvec": nvec, "grid": [np0, np1, np2], "nst_per_gpu": nst, "nprna": sum(s["na"] * s["npr"] for s in species),
                     "fused_plane_path": ft.fused(), "states_per_batch": ft.batch()}}
    del ft, nlp, c, v, hpsi, rho
    torch.cuda.empty_cache()
    return out


if __name__ == "__main__":
    main()
