#!/usr/bin/env python3
"""One MSM of 2^LOG_N points on resident bases and scalars (for ncu captures of the MSM kernels)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import snarkos_b200 as S
S.init(0)
n = 1 << int(os.environ.get("LOG_N", "24"))
bases = S.synthetic_bases(n, seed=5)
g = torch.Generator(device="cuda"); g.manual_seed(1)
sc = torch.randint(-(1 << 63), (1 << 63) - 1, (n, 4), dtype=torch.int64, device="cuda", generator=g)
sc[:, 3] &= (1 << 60) - 1
rb = S.ResidentBases(bases)   # stored as (P, phi(P)) pairs: the GLV path, as VariableBase.msm takes
for _ in range(int(os.environ.get("CALLS", "1"))):
    rb.msm(sc)
torch.cuda.synchronize()
