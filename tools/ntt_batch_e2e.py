#!/usr/bin/env python3
"""Host-buffer NTT of a batch (BASELINE config 2: 2^20 x 16) with and without the three-stream group pipeline."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import snarkos_b200 as S
S.init(0)
n, batch = 1 << 20, 16
d = S.EvaluationDomain(n)
t = torch.randint(0, 1 << 59, (batch, n, 4), dtype=torch.int64).pin_memory()
for mode in ("pipelined", "single"):
    if mode == "single":
        S.set_option("ntt_host_pipeline", 0)
    d.fft_in_place(t)
    t0 = time.perf_counter()
    for _ in range(5):
        d.fft_in_place(t)
    ms = (time.perf_counter() - t0) / 5 * 1e3
    print(f"2^20 x 16 host NTT {mode}: {ms:.2f} ms, {batch * n / ms / 1e6:.2f} Gelem/s", flush=True)
