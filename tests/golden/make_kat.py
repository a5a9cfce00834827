#!/usr/bin/env python3
"""Extract the offline-derived known-answer vectors of SURVEY.md appendix A.3 / section 8c into
tests/golden/kat.json.  SURVEY.md derived them from first principles with Python big-ints (not
from snarkVM, whose source is absent), so they pin the oracle, and the oracle pins the GPU.
Run from the repo root:  python tests/golden/make_kat.py
"""
import json, os, re

root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
txt = open(os.path.join(root, "SURVEY.md")).read()
a3 = txt[txt.index("### A.3 Known-answer vectors"):txt.index("### A.4")]


def ints(s):
    return [int(x) for x in re.findall(r"(?<![\w^])\d{20,}(?!\w)", s)]


def block(start, end):
    return a3[a3.index(start):a3.index(end)]


kat = {}
g = ints(block("G  =", "2G ="));   kat["G"] = g[:2]
g = ints(block("2G =", "3G ="));   kat["2G"] = g[:2]
g = ints(block("3G =", "MSM("));   kat["3G"] = g[:2]
g = ints(block("MSM(", "(r-1)*G")); kat["MSM_1_2_3__G_2G_3G"] = g[:2]
for name in ("omega_4", "omega_8", "omega_2^20", "(2^20)^-1", "22^-1"):
    m = re.search(re.escape(name) + r"\s*=\s*(\d+)", a3)
    kat[name] = int(m.group(1))
kat["NTT_4_1234"] = [10] + ints(block("NTT_4([1,2,3,4])", "cosetNTT_4"))
kat["cosetNTT_4_1234"] = [44089] + ints(block("cosetNTT_4([1,2,3,4])", "NTT_8("))
kat["NTT_8_1to8"] = [36] + ints(block("NTT_8([1..8])", "```\n\n"))
# section 8c constants (limb form)
c8 = txt[txt.index("Golden/known-answer material available"):txt.index("Oracle self-checks to implement")]
kat["r_hex"] = re.search(r"\*\*r\*\* = x⁴−x²\+1 = `(0x[0-9a-f]+)`", c8).group(1)
kat["p_hex"] = re.search(r"\*\*p\*\* = .*? = `(0x[0-9a-f]+)`", c8).group(1)
kat["fr_R_limbs"] = [int(x) for x in re.search(r"R=2²⁵⁶ mod r `\[([^\]]+)\]`", c8).group(1).split(",")]
kat["fr_R2_limbs"] = [int(x) for x in re.search(r"R² `\[(2726216793283724667[^\]]+)\]`", c8).group(1).split(",")]
kat["fr_gen22_mont_limbs"] = [int(x) for x in re.search(r"g = 22 \(Mont limbs `\[([^\]]+)\]`", c8).group(1).split(",")]
kat["fr_two_adic_root"] = int(re.search(r"22\^\(\(r−1\)/2⁴⁷\) = `(\d+)`", c8).group(1))
kat["fr_two_adic_root_mont_limbs"] = [int(x) for x in re.search(r"\(Mont limbs `\[(12646347781564978760[^\]]+)\]`", c8).group(1).split(",")]
kat["fq_R_limbs"] = [int(x) for x in re.search(r"R=2³⁸⁴ mod p `\[([^\]]+)\]`", c8).group(1).split(",")]
kat["fq_R2_limbs_hex"] = [x.strip() for x in re.search(r"R² `\[(0xb786686c9400cd22[^\]]+)\]`", c8).group(1).split(",")]
assert len(kat["NTT_4_1234"]) == 4 and len(kat["cosetNTT_4_1234"]) == 4 and len(kat["NTT_8_1to8"]) == 8
out = os.path.join(root, "tests/golden/kat.json")
json.dump({k: (str(v) if isinstance(v, int) else [str(x) for x in v] if isinstance(v, list) else v)
           for k, v in kat.items()}, open(out, "w"), indent=1)
print("wrote", out, len(kat), "entries")
