"""times qb200_diag on a synthetic Hermitian subspace problem of n states (tools: run under ncu for per-kernel durations)"""
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "oracle"))
from qball_b200 import basis as BB, host as H, synth as S

n = int(sys.argv[1]) if len(sys.argv) > 1 else 768
cell, ecut = (16, 0, 0, 0, 16, 0, 0, 0, 16), 8.0
b = BB.make_basis(cell, ecut, (0.1, 0.2, 0.3), False)
c = S.synth_coefficients(b["kpg2"], ecut, n, b["ngw"], False, 5)
hc = S.synth_coefficients(b["kpg2"], ecut, n, b["ngw"], False, 6) + c * (1.0 + b["kpg2"])[None, :]
st = torch.cuda.Stream()
la = H.SubspaceLA(b, stream=st)
with torch.cuda.stream(st):
    cd, hd = torch.from_numpy(c).cuda(), torch.from_numpy(hc).cuda()
    la.gram(cd)
    for rep in range(int(os.environ.get("REPS", "3"))):
        c2 = cd.clone()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        w, sw = la.diag(c2, hd)
        torch.cuda.synchronize()
        print(f"n {n} ngw {b['ngw']} diag {1e3 * (time.perf_counter() - t0):.2f} ms, {sw} sweeps, blocked {la.query(14)}, w[0] {w[0]:.6f}", flush=True)
