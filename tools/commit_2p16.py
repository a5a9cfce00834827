#!/usr/bin/env python3
"""One KZG commit against 2^16 resident powers (the size a Varuna proof of transfer_public commits against), for ncu:
warm-up calls, then ONE call between cudaProfilerStart / Stop."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import snarkos_b200 as S
S.init(0)
n = 1 << int(os.environ.get("LOG_N", "16"))
powers = S.Powers(S.synthetic_bases(n, seed=5))
g = torch.Generator(device="cuda"); g.manual_seed(2)
c = torch.randint(-(1 << 63), (1 << 63) - 1, (n, 4), dtype=torch.int64, device="cuda", generator=g)
c[:, 3] &= (1 << 60) - 1
for _ in range(3):
    S.KZG10.commit(powers, c)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    S.KZG10.commit(powers, c)
e1.record(); torch.cuda.synchronize()
print(f"commit 2^{n.bit_length() - 1}: {e0.elapsed_time(e1) / 10:.3f} ms per call", flush=True)
torch.cuda.cudart().cudaProfilerStart()
S.KZG10.commit(powers, c)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
with S.profile() as p:
    S.KZG10.commit(powers, c)
print("stages: " + " ".join(f"{k}={v:.3f}" for k, v in p.totals().items()), flush=True)
