import sys
sys.path.insert(0,'/root/repo')
from oracle import bls12_377 as O
x = 0x8508c00000000001
r, p = O.R_MOD, O.P_MOD
lam = x*x - 1
assert (lam*lam + lam + 1) % r == 0
print("lambda bits", lam.bit_length(), hex(lam))
# beta: primitive cube root of unity in Fq with phi(P) = (beta x, y) = lam * P
g = 2
while True:
    b = pow(g, (p-1)//3, p)
    if b != 1: break
    g += 1
G = O.G1_GEN
lamG = O.g1_mul(G, lam)
for beta in (b, b*b % p):
    if (beta*G[0] % p, G[1]) == lamG:
        print("beta ok", hex(beta)); BETA = beta
# barrett
SH = 254
m = (1 << SH) // lam
print("m bits", m.bit_length(), hex(m))
import random
random.seed(1)
mx = 0
for _ in range(200000):
    k = random.randrange(r)
    q = (k*m) >> SH
    rem = k - q*lam
    fix = 0
    while rem >= lam: rem -= lam; q += 1; fix += 1
    assert rem >= 0 and q*lam + rem == k and q < (1<<127) and rem < (1<<127)
    mx = max(mx, fix)
print("max fixups", mx)
for k in (0, 1, r-1, lam, lam-1, lam+1, 2*lam, r-2):
    q = (k*m) >> SH; rem = k - q*lam
    while rem >= lam: rem -= lam; q += 1
    assert rem >= 0 and q*lam+rem == k and q < (1<<127)
R384 = 1 << 384
print("beta mont limbs", [hex((BETA*R384 % p >> (32*i)) & 0xffffffff) for i in range(12)])
print("lam limbs", [hex((lam >> (32*i)) & 0xffffffff) for i in range(4)])
print("m limbs", [hex((m >> (32*i)) & 0xffffffff) for i in range(5)])
