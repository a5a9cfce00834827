#!/usr/bin/env python3
"""Small / medium MSMs against resident bases: plain vs window-table registration (no fold), device stage sums and
host-observed latency per call."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import snarkos_b200 as S
S.init(0)
for log_n in (14, 16, 18, 20):
    n = 1 << log_n
    bases = S.synthetic_bases(n, seed=5)
    g = torch.Generator(device="cuda"); g.manual_seed(1)
    sc = torch.randint(-(1 << 63), (1 << 63) - 1, (n, 4), dtype=torch.int64, device="cuda", generator=g)
    sc[:, 3] &= (1 << 60) - 1
    hs = sc.cpu().pin_memory()
    for name, rb in (("plain", S.ResidentBases(bases)), ("table", S.ResidentBases(bases, tabulate=True))):
        rb.msm(sc); torch.cuda.synchronize()
        with S.profile() as p:
            rb.msm(sc)
        tot = sum(v for _, v in p.stages)
        t0 = time.perf_counter()
        for _ in range(10):
            rb.msm(hs)                                     # host scalars in, 144 B out: what a KZG commit call sees
        wall = (time.perf_counter() - t0) / 10 * 1e3
        print(f"2^{log_n} {name}: device {tot:.2f} ms, host-call {wall:.2f} ms  " + " ".join(f"{k[4:]}={v:.2f}" for k, v in p.totals().items()), flush=True)
        rb.release()

# chunk length of the XYZZ walk on a small tabulated commit, and k commits in one launch set
import numpy as np
for log_n in (14, 16):
    n = 1 << log_n
    bases = S.synthetic_bases(n, seed=5)
    powers = S.Powers(bases)
    g = torch.Generator(device="cuda"); g.manual_seed(2)
    polys = []
    for _ in range(8):
        c = torch.randint(-(1 << 63), (1 << 63) - 1, (n, 4), dtype=torch.int64, device="cuda", generator=g)
        c[:, 3] &= (1 << 60) - 1
        polys.append(c)
    for chunk in (0, 8, 16, 32, 64):
        S.set_option("msm_chunk", chunk)
        S.KZG10.commit(powers, polys[0]); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            S.KZG10.commit(powers, polys[0])
        e1.record(); torch.cuda.synchronize()
        one = e0.elapsed_time(e1) / 10
        S.KZG10.commit_batch(powers, polys); torch.cuda.synchronize()
        e0.record()
        for _ in range(10):
            S.KZG10.commit_batch(powers, polys)
        e1.record(); torch.cuda.synchronize()
        print(f"2^{log_n} commit chunk={chunk}: one {one:.3f} ms, 8 in one launch set {e0.elapsed_time(e1) / 10:.3f} ms", flush=True)
    S.set_option("msm_chunk", 0)
    with S.profile() as p:
        S.KZG10.commit_batch(powers, polys)
    print(f"2^{log_n} commit_batch x8 stages: " + " ".join(f"{k[4:]}={v:.2f}" for k, v in p.totals().items()), flush=True)
    powers.release()
