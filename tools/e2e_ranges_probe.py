#!/usr/bin/env python3
"""Host-buffer MSM 2^24 (pinned) end to end over the schedule of the streamed ranges: first range 2^first_log points,
doubling up to 2^chunk_log (options msm_host_first_log / msm_host_chunk_log)."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import snarkos_b200 as S
S.init(0)
n = 1 << int(os.environ.get("LOG_N", "24"))
bases = S.synthetic_bases(n, seed=5)
g = torch.Generator(device="cuda"); g.manual_seed(1)
sc = torch.randint(-(1 << 63), (1 << 63) - 1, (n, 4), dtype=torch.int64, device="cuda", generator=g)
sc[:, 3] &= (1 << 60) - 1
hb, hs = bases.cpu().pin_memory(), sc.cpu().pin_memory()
ref = None
for chunk_log in (23, 22):
    for first_log in (17, 18, 19, 20, 21):
        S.set_option("msm_host_first_log", first_log)
        S.set_option("msm_host_chunk_log", chunk_log)
        out = S.VariableBase.msm(hb, hs)
        t0 = time.perf_counter()
        for _ in range(4):
            out = S.VariableBase.msm(hb, hs)
        ms = (time.perf_counter() - t0) / 4 * 1e3
        comp = S.g1_compress(out.reshape(1, 144)).tobytes()
        ref = ref or comp
        print(json.dumps({"first_log": first_log, "chunk_log": chunk_log, "ms": round(ms, 2), "mpoints_s": round(n / ms / 1e3, 1),
                          "same_point": comp == ref}), flush=True)
