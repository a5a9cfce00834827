#!/usr/bin/env python3
"""Bucket reduction of one MSM over the segment length and the lane layout (one lane / a quad of lanes per segment):
stage events of b200_msm_g1_bls12_377_device at 2^LOG_N, result compared with the default configuration's."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import snarkos_b200 as S
S.init(0)
n = 1 << int(os.environ.get("LOG_N", "24"))
bases = S.synthetic_bases(n, seed=5)
g = torch.Generator(device="cuda"); g.manual_seed(1)
sc = torch.randint(-(1 << 63), (1 << 63) - 1, (n, 4), dtype=torch.int64, device="cuda", generator=g)
sc[:, 3] &= (1 << 60) - 1


def run(seg, quad_max):
    S.set_option("msm_seg_len", seg)
    S.set_option("msm_reduce_quad_max", quad_max)
    for _ in range(2):
        out = S.VariableBase.msm(bases, sc)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        out = S.VariableBase.msm(bases, sc)
    e1.record(); torch.cuda.synchronize()
    with S.profile() as p:
        S.VariableBase.msm(bases, sc)
    t = p.totals()
    return out, {"seg_len": seg, "quad_max": quad_max, "ms": round(e0.elapsed_time(e1) / 3, 3),
                 "reduce_segments": round(t.get("msm_reduce_segments", 0), 3), "window_sum": round(t.get("msm_window_sum", 0), 3),
                 "fold": round(t.get("msm_fold", 0), 3)}


ref, r = run(0, 8192)
print(json.dumps(r), flush=True)
ref = S.g1_compress(ref.reshape(1, 144)).cpu()
for quad_max in (8192, 1 << 30):
    for seg in (8, 16, 32, 48, 64, 96, 128):
        out, r = run(seg, quad_max)
        r["same_point"] = bool(torch.equal(S.g1_compress(out.reshape(1, 144)).cpu(), ref))
        print(json.dumps(r), flush=True)
