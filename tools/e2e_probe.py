#!/usr/bin/env python3
"""Host-buffer MSM 2^24 (pinned) end to end for several window widths of the streamed ranges."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import snarkos_b200 as S
S.init(0)
n = 1 << 24
bases = S.synthetic_bases(n, seed=5)
g = torch.Generator(device="cuda"); g.manual_seed(1)
sc = torch.randint(-(1 << 63), (1 << 63) - 1, (n, 4), dtype=torch.int64, device="cuda", generator=g)
sc[:, 3] &= (1 << 60) - 1
hb, hs = bases.cpu().pin_memory(), sc.cpu().pin_memory()
for c in [int(x) for x in os.environ.get("CS", "0,17,18,19").split(",")]:
    if c:
        S.set_option("msm_window_bits", c)
    S.VariableBase.msm(hb, hs)
    t0 = time.perf_counter()
    for _ in range(3):
        S.VariableBase.msm(hb, hs)
    ms = (time.perf_counter() - t0) / 3 * 1e3
    print(f"e2e 2^24 c={c}: {ms:.1f} ms  {n / ms / 1e3:.1f} Mpoints/s", flush=True)
