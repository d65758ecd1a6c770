"""one qb200_residual + qb200_gram + qb200_compute_current on an MgO216-size block, for ncu captures (tools/gpu_round.sh)"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from qball_b200 import basis as B, host as H, synth as S  # noqa: E402

cell, ecut = (23.1, 0, 0, 0, 23.1, 0, 0, 0, 23.1), 25.0
b = B.make_basis(cell, ecut, (0.0, 0.0, 0.0), True)
nst = int(sys.argv[1]) if len(sys.argv) > 1 else 768
c = torch.from_numpy(S.synth_coefficients(b["kpg2"], ecut, nst, b["ngw"], False, seed=1)).cuda()
hc = torch.from_numpy(S.synth_coefficients(b["kpg2"], ecut, nst, b["ngw"], False, seed=2)).cuda()
la = H.SubspaceLA(b)
for _ in range(2):
    la.residual(c, hc)
    cg = c.clone()
    la.gram(cg)
torch.cuda.synchronize()
g = cg.cpu().numpy()[:8]
print("orthonormality of the first 8 states:", np.abs(g.conj() @ g.T - np.eye(8)).max(), "launches", la.launches())
