#!/usr/bin/env python3
"""NTT kernel variants side by side: default pass kernel (per-thread loads, planar tile) vs the bulk-copy (TMA,
cp.async.bulk + mbarrier, persistent double-buffered CTAs) variant, for several tile sizes.  Prints ms per transform;
every configuration must produce the same bytes as the first one."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import snarkos_b200 as S

S.init(0)
out = []
CONFIGS = ((0, 11, ""), (0, 10, ""), (1, 10, ""), (1, 11, ""), (1, 9, ""))
if os.environ.get("NTT_CONFIGS"):            # "variant:tile_log:plan;..." e.g. "0:11:;2:10:;2:10:8,8,8"
    CONFIGS = tuple((int(a), int(b), c) for a, b, c in (x.split(":") for x in os.environ["NTT_CONFIGS"].split(";")))
SIZES = ((24, 1), (20, 16), (16, 64), (26, 1))
if os.environ.get("NTT_SIZES"):              # "log_n:batch;..."
    SIZES = tuple((int(a), int(b)) for a, b in (x.split(":") for x in os.environ["NTT_SIZES"].split(";")))
for log_n, batch in SIZES:
    n = 1 << log_n
    gen = torch.Generator(device="cuda"); gen.manual_seed(log_n)
    src = torch.randint(0, 1 << 59, (batch, n, 4), dtype=torch.int64, device="cuda", generator=gen)
    d = S.EvaluationDomain(n)
    ref = None
    for variant, tile_log, plan in CONFIGS:
        S.set_option("ntt_variant", variant); S.set_option("ntt_tile_log", tile_log); S.set_option("ntt_plan", plan)
        try:
            x = src.clone()
            d.fft_in_place(x)
            torch.cuda.synchronize()
            if ref is None:
                ref = x.clone()
            same = bool(torch.equal(x, ref))
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 10
            e0.record()
            for _ in range(reps):
                d.fft_in_place(x)
            e1.record()
            torch.cuda.synchronize()
            with S.profile() as prof:
                d.fft_in_place(x)
            rec = {"log_n": log_n, "batch": batch, "variant": variant, "tile_log": tile_log, "plan": plan, "ms": e0.elapsed_time(e1) / reps,
                   "gelem_s": n * batch / (e0.elapsed_time(e1) / reps * 1e-3) / 1e9, "same_bytes": same, "stages": prof.totals()}
        except Exception as ex:
            rec = {"log_n": log_n, "batch": batch, "variant": variant, "tile_log": tile_log, "error": repr(ex)}
        print(json.dumps(rec), flush=True)
        out.append(rec)
    del src, ref
    torch.cuda.empty_cache()
S.set_option("ntt_variant", 0); S.set_option("ntt_tile_log", 11)
