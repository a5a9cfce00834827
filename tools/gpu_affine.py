#!/usr/bin/env python3
"""Stage timings of the MSM for several counts of batched-affine rounds and window widths.
Env: LOG_N (24), ROUNDS ("0,3,4,5"), CS ("20")."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import snarkos_b200 as S
S.init(0)
log_n = int(os.environ.get("LOG_N", "24"))
n = 1 << log_n
bases = S.synthetic_bases(n, seed=5)
g = torch.Generator(device="cuda"); g.manual_seed(1)
sc = torch.randint(-(1 << 63), (1 << 63) - 1, (n, 4), dtype=torch.int64, device="cuda", generator=g)
sc[:, 3] &= (1 << 60) - 1
rb = S.ResidentBases(bases)
plain = os.environ.get("PLAIN") == "1"            # VariableBase.msm on raw points (GLV path) instead of resident bases
call = (lambda: S.VariableBase.msm(bases, sc)) if plain else (lambda: rb.msm(sc))
ref = None
for c in [int(x) for x in os.environ.get("CS", "20").split(",")]:
    for r in [int(x) for x in os.environ.get("ROUNDS", "0,3,4,5").split(",")]:
        if c > 0:
            S.set_option("msm_window_bits", c)
        if r >= 0:
            S.set_option("msm_affine_rounds", r)
        out = call(); torch.cuda.synchronize()
        best = None
        for _ in range(2):
            with S.profile() as p:
                call()
            tot = sum(v for _, v in p.stages)
            if best is None or tot < best[0]:
                best = (tot, p.totals())
        print(f"2^{log_n} c={c} rounds={r}: total {best[0]:.2f} ms  " + " ".join(f"{k[4:]}={v:.2f}" for k, v in best[1].items()), flush=True)
