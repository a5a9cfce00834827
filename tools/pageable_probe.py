import os, sys, time
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import snarkos_b200 as S
S.init(0)
n = 1 << 24
bases = S.synthetic_bases(n, seed=5)
g = torch.Generator(device="cuda"); g.manual_seed(1)
sc = torch.randint(-(1 << 63), (1 << 63) - 1, (n, 4), dtype=torch.int64, device="cuda", generator=g)
sc[:, 3] &= (1 << 60) - 1
hb_pin = bases.cpu().pin_memory(); hs_pin = sc.cpu().pin_memory()
hb = bases.cpu().numpy().copy(); hs = sc.cpu().numpy().copy()
rb = S.ResidentBases(bases)
for name, fn in (("pinned full", lambda: S.VariableBase.msm(hb_pin, hs_pin)), ("pageable full", lambda: S.VariableBase.msm(hb, hs.view(np.uint64))),
                 ("pinned resident", lambda: rb.msm(hs_pin)), ("pageable resident", lambda: rb.msm(hs.view(np.uint64)))):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3): fn()
    torch.cuda.synchronize()
    print(name, (time.perf_counter() - t0) / 3 * 1e3, "ms", flush=True)
