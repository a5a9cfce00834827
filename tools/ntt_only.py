#!/usr/bin/env python3
"""One forward NTT of 2^24 elements on resident data (for ncu captures of ntt_pass_kernel)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import snarkos_b200 as S
S.init(0)
n = 1 << int(os.environ.get("LOG_N", "24"))
t = torch.randint(0, 1 << 59, (n, 4), dtype=torch.int64, device="cuda")
d = S.EvaluationDomain(n)
for _ in range(2):
    d.fft_in_place(t)
torch.cuda.synchronize()
print("ok")
