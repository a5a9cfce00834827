import os, sys
sys.path.insert(0, "/root/repo")
import torch
import snarkos_b200 as S
S.init(0)
n = 1 << int(os.environ.get("LOG_N", "24"))
bases = S.synthetic_bases(n, seed=5)
g = torch.Generator(device="cuda"); g.manual_seed(1)
sc = torch.randint(-(1 << 63), (1 << 63) - 1, (n, 4), dtype=torch.int64, device="cuda", generator=g)
sc[:, 3] &= (1 << 60) - 1
for _ in range(2):
    S.VariableBase.msm(bases, sc)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    S.VariableBase.msm(bases, sc)
e1.record(); torch.cuda.synchronize()
with S.profile() as p:
    S.VariableBase.msm(bases, sc)
print("msm 2^%s: %.2f ms" % (os.environ.get("LOG_N", "24"), e0.elapsed_time(e1) / 3), {k[4:]: round(v, 2) for k, v in p.totals().items()}, "xyzz fallbacks:", S.counter("msm_xyzz_fallbacks"))
