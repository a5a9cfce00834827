#!/usr/bin/env python3
"""2^24 MSM with 4 .. 8 batched-affine pair rounds before the XYZZ finish (option msm_affine_rounds)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import snarkos_b200 as S
S.init(0)
n = 1 << 24
bases = S.synthetic_bases(n, seed=5)
g = torch.Generator(device="cuda"); g.manual_seed(1)
sc = torch.randint(-(1 << 63), (1 << 63) - 1, (n, 4), dtype=torch.int64, device="cuda", generator=g)
sc[:, 3] &= (1 << 60) - 1
ref = None
for r in (5, 4, 6, 7, 8):
    S.set_option("msm_affine_rounds", r)
    out = S.VariableBase.msm(bases, sc); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        S.VariableBase.msm(bases, sc)
    e1.record(); torch.cuda.synchronize()
    with S.profile() as p:
        S.VariableBase.msm(bases, sc)
    t = p.totals()
    print(f"rounds={r}: {e0.elapsed_time(e1) / 3:.2f} ms  accumulate={t.get('msm_accumulate', 0):.2f} pairs_add={t.get('msm_pairs_add', 0):.2f} "
          f"pairs_denoms={t.get('msm_pairs_denoms', 0):.2f} pairs_invert={t.get('msm_pairs_invert', 0):.2f} reduce={t.get('msm_reduce_segments', 0):.2f}", flush=True)
S.set_option("msm_affine_rounds", -1)
