import os, sys
sys.path.insert(0, "/root/repo")
import torch
import snarkos_b200 as S
S.init(0)
def mk(log_n, seed):
    n = 1 << log_n
    bases = S.synthetic_bases(n, seed=seed)
    g = torch.Generator(device="cuda"); g.manual_seed(seed)
    sc = torch.randint(-(1 << 63), (1 << 63) - 1, (n, 4), dtype=torch.int64, device="cuda", generator=g)
    sc[:, 3] &= (1 << 60) - 1
    return bases, sc
def timeit(b, s, tag):
    S.VariableBase.msm(b, s); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        S.VariableBase.msm(b, s)
    e1.record(); torch.cuda.synchronize()
    with S.profile() as p:
        S.VariableBase.msm(b, s)
    print(tag, "%.2f ms" % (e0.elapsed_time(e1) / 3), "fallbacks", S.counter("msm_xyzz_fallbacks"), {k[4:]: round(v, 1) for k, v in p.totals().items() if v > 3}, flush=True)
b24, s24 = mk(24, 1)
timeit(b24, s24, "2^24 first")
mode = os.environ.get("MODE", "release")
if mode == "release":
    S.release_scratch(); torch.cuda.empty_cache()
b25, s25 = mk(25, 2)
timeit(b25, s25, "2^25 after 2^24 (%s)" % mode)
timeit(b24, s24, "2^24 again")
print(torch.cuda.mem_get_info())
