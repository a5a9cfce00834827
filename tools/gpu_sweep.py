#!/usr/bin/env python3
"""Per-stage device timings (CUDA events via b200_profile_*) for MSM window widths and NTT pass plans.
Writes gpurun_out/sweep.json.  Usage: python tools/gpu_sweep.py [msm|ntt|micro ...]"""
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import snarkos_b200 as S

what = set(sys.argv[1:]) or {"msm", "ntt", "micro"}
torch.cuda.set_device(0)
S.init(0)
dev = torch.device("cuda", 0)
res = {}


def rand_limbs(count, seed=0):
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    t = torch.randint(-(1 << 63), (1 << 63) - 1, (count, 4), dtype=torch.int64, device=dev, generator=g)
    t[:, 3] &= (1 << 60) - 1
    return t


def staged(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    best = None
    for _ in range(reps):
        with S.profile() as p:
            fn()
        tot = sum(v for _, v in p.stages)
        if best is None or tot < best[0]:
            best = (tot, p.totals())
    return best


if "micro" in what:
    names = {0: "imad_lo", 2: "imad_hi", 7: "imad_wide_noaddend", 1: "imad_wide+iadd3x2(ptxas split)", 14: "imad_wide_addend",
             8: "imad_wide_x_chain", 6: "iadd3", 11: "dfma", 3: "fr_modmul", 4: "fq_modmul", 5: "xyzz_madd"}
    iters = {0: 8192, 1: 8192, 2: 8192, 6: 8192, 7: 8192, 8: 2048, 11: 8192, 14: 8192, 3: 512, 4: 256, 5: 64}
    res["micro"] = {}
    for kind, name in names.items():
        ms, ops = ctypes.c_float(), ctypes.c_double()
        best = 0
        for _ in range(3):
            S._lib.check(S.lib().b200_debug_microbench(kind, iters[kind], ctypes.byref(ms), ctypes.byref(ops)))
            best = max(best, ops.value / (ms.value * 1e-3))
        res["micro"][name] = best
        print(f"micro {name:32s} {best / 1e9:10.2f} Gop/s  ({best / 148 / 1.965e9:6.2f} per SM per clk @1965MHz)", flush=True)

if "msm" in what:
    res["msm"] = {}
    chunks = [int(x) for x in os.environ.get("SWEEP_CHUNKS", "64").split(",")]
    for log_n, cs in ((16, (10, 11, 12, 13, 14)), (20, (14, 15, 16, 17, 18)), (24, (17, 18, 19, 20, 21))):
        n = 1 << log_n
        bases = S.synthetic_bases(n, seed=5)
        sc = rand_limbs(n, 1)
        rb = S.ResidentBases(bases)
        for c in cs:
            for ch in chunks:
                S.set_option("msm_window_bits", c)
                S.set_option("msm_chunk", ch)
                tot, st = staged(lambda: rb.msm(sc), reps=2)
                res["msm"][f"2^{log_n} c={c} chunk={ch}"] = {"total_ms": tot, **st}
                print(f"msm 2^{log_n} c={c} chunk={ch}: total {tot:8.3f} ms  " + " ".join(f"{k[4:]}={v:.3f}" for k, v in st.items()), flush=True)
        S.set_option("msm_window_bits", 0)
        S.set_option("msm_chunk", 0)
        rb.release()
        del bases, sc

if "ntt" in what:
    res["ntt"] = {}
    cases = [
        (16, 16, ["8,8", "6,5,5"], [11, 12]),
        (20, 16, ["10,10", "7,7,6", "9,11", "8,12"], [10, 11, 12]),
        (24, 1, ["12,12", "8,8,8", "9,9,6", "10,10,4", "11,11,2"], [10, 11, 12]),
    ]
    for log_n, batch, plans, tiles in cases:
        n = 1 << log_n
        t = rand_limbs(batch * n, 2)
        d = S.EvaluationDomain(n)
        for plan in plans:
            for tile in tiles:
                lens = [int(x) for x in plan.split(",")]
                if max(lens) > tile:
                    continue
                S.set_option("ntt_plan", plan)
                S.set_option("ntt_tile_log", tile)
                tot, st = staged(lambda: d.fft_in_place(t))
                key = f"2^{log_n}x{batch} plan={plan} tile=2^{tile}"
                res["ntt"][key] = {"total_ms": tot, **st}
                print(f"ntt {key}: total {tot:8.3f} ms  " + " ".join(f"{k[4:]}={v:.3f}" for k, v in st.items()), flush=True)
        S.set_option("ntt_plan", "")
        S.set_option("ntt_tile_log", 11)
        del t

os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "sweep.json"), "w"), indent=1)
