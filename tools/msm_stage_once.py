#!/usr/bin/env python3
"""2^24 device-resident MSM: wall time (5 calls) and stage events of one call."""
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
S.VariableBase.msm(bases, sc); torch.cuda.synchronize()
ts = []
for _ in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); S.VariableBase.msm(bases, sc); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
with S.profile() as p:
    S.VariableBase.msm(bases, sc)
print(f"2^{n.bit_length() - 1}: min {min(ts):.2f} med {sorted(ts)[2]:.2f} ms  " + " ".join(f"{k[4:]}={v:.2f}" for k, v in p.totals().items()), flush=True)
