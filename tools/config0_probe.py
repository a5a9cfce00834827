#!/usr/bin/env python3
"""BASELINE configs[0]: VariableBase::msm of 2^16 random bases / scalars from host buffers (and smaller / larger sizes)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import snarkos_b200 as S
from oracle import c_oracle as C
S.init(0)
for log_n in (12, 14, 16, 18, 20):
    n = 1 << log_n
    bases = S.synthetic_bases(n, seed=5)
    g = torch.Generator(device="cuda"); g.manual_seed(1)
    sc = torch.randint(-(1 << 63), (1 << 63) - 1, (n, 4), dtype=torch.int64, device="cuda", generator=g)
    sc[:, 3] &= (1 << 60) - 1
    hb, hs = bases.cpu().numpy(), sc.cpu().numpy().view(np.uint64)           # pageable, as a Rust Vec is
    S.VariableBase.msm(hb, hs)
    reps = 10
    t0 = time.perf_counter()
    for _ in range(reps):
        out = S.VariableBase.msm(hb, hs)
    ms = (time.perf_counter() - t0) / reps * 1e3
    t0 = time.perf_counter()
    ref = C.msm(hb, hs, nthreads=len(os.sched_getaffinity(0)))
    cpu_ms = (time.perf_counter() - t0) * 1e3
    print(f"2^{log_n}: GPU host-call {ms:.2f} ms ({n / ms / 1e3:.2f} Mpoints/s), CPU port {cpu_ms:.1f} ms ({n / cpu_ms / 1e3:.3f} Mpoints/s), x{cpu_ms / ms:.0f}", flush=True)
