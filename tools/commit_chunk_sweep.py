#!/usr/bin/env python3
"""KZG commit against resident powers over the chunk length of the XYZZ walk (entries per thread): the walk of a small
call is one dependent chain per thread on a partly idle chip, so what matters is whether the grid is one wave."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import snarkos_b200 as S
S.init(0)
for log_n in (14, 15, 16, 17, 18):
    n = 1 << log_n
    powers = S.Powers(S.synthetic_bases(n, seed=5))
    g = torch.Generator(device="cuda"); g.manual_seed(2)
    c = torch.randint(-(1 << 63), (1 << 63) - 1, (n, 4), dtype=torch.int64, device="cuda", generator=g)
    c[:, 3] &= (1 << 60) - 1
    ref = None
    for chunk in (0, 8, 10, 12, 14, 16, 18, 20, 24, 28, 32, 40, 48, 64):
        S.set_option("msm_chunk", chunk)
        out = S.KZG10.commit(powers, c); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            out = S.KZG10.commit(powers, c)
        e1.record(); torch.cuda.synchronize()
        with S.profile() as p:
            S.KZG10.commit(powers, c)
        t = p.totals()
        comp = bytes(S.g1_compress(torch.as_tensor(out).reshape(1, 144)).cpu().numpy().tobytes()) if isinstance(out, torch.Tensor) else None
        ref = ref or comp
        print(json.dumps({"log_n": log_n, "chunk": chunk, "ms": round(e0.elapsed_time(e1) / 20, 4), "accumulate": round(t.get("msm_accumulate", 0), 3),
                          "combine": round(t.get("msm_combine", 0), 3), "same_point": comp == ref}), flush=True)
    S.set_option("msm_chunk", 0)
    powers.release()
