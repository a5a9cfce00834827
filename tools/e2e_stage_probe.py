#!/usr/bin/env python3
"""Host-buffer MSM 2^24 (pinned): wall time and the library's stage events for several first-range sizes."""
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
TWO = [int(x) for x in os.environ.get("TWO", "1,0").split(",")]
for two, fl in [(t, f) for t in TWO for f in [int(x) for x in os.environ.get("FIRST", "19,20,21").split(",")]]:
    S.set_option("msm_host_first_log", fl)
    S.set_option("msm_stream_two", two)
    out = S.VariableBase.msm(hb, hs)
    ts = []
    for _ in range(5):
        t0 = time.perf_counter()
        S.VariableBase.msm(hb, hs)
        ts.append((time.perf_counter() - t0) * 1e3)
    with S.profile() as p:
        S.VariableBase.msm(hb, hs)
    t = p.totals()
    print(f"two={two} first_log={fl}: min {min(ts):.1f} med {sorted(ts)[2]:.1f} max {max(ts):.1f} ms  stages(sum {sum(t.values()):.1f}): " +
          " ".join(f"{k[4:]}={v:.2f}" for k, v in t.items()), flush=True)
