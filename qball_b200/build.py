"""Builds libqball_b200.so (hand-written CUDA for sm_100a + the C ABI) in-tree with nvcc: one object per source, compiled
in parallel, then linked."""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
OBJ = os.path.join(PKG, "build")
LIB = os.path.join(PKG, "libqball_b200.so")
SOURCES = ["transform.cu", "plane.cu", "zcol_tmem.cu", "ycols_tmem.cu", "nonlocal.cu", "hpsi.cu", "diag.cu", "comm.cu"]
NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-Xcompiler", "-fPIC",
              "-Xptxas", "-v"]


def _stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(os.path.dirname(PKG), "include", "qball_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not (force or _stale()):
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    os.makedirs(OBJ, exist_ok=True)
    srcs = [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]

    def compile_one(s):
        o = os.path.join(OBJ, s.replace(".cu", ".o"))
        cmd = [nvcc, "-c", "-o", o] + NVCC_FLAGS + [os.path.join(CSRC, s)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        return s, o, cmd, r

    with ThreadPoolExecutor(max_workers=len(srcs)) as ex:
        results = list(ex.map(compile_one, srcs))
    log = os.path.join(PKG, "build.log")
    failed = False
    with open(log, "w") as f:
        for s, o, cmd, r in results:
            f.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
            failed = failed or r.returncode != 0
            if verbose or r.returncode:
                sys.stderr.write(r.stdout + r.stderr)
    if failed:
        raise RuntimeError(f"nvcc failed; see {log}")
    cmd = [nvcc, "-shared", "-o", LIB] + [o for _, o, _, _ in results] + ["-lcudart", "-ldl"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    with open(log, "a") as f:
        f.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
    if r.returncode:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError(f"link failed; see {log}")
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose="-v" in sys.argv))
