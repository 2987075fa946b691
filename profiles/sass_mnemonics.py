"""Blackwell-specific SASS mnemonics per kernel of libodb200.so (runs on the CPU box: cuobjdump -sass).

    python profiles/sass_mnemonics.py > profiles/r2_sass_mnemonics.txt
"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "opendiloco_b200", "_C", "libodb200.so")
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
KEEP = re.compile(r"^(UTC|UTMA|UBLKCP|LDTM|STTM|LDGMC|STGMC|REDGMC|MULTIMEM|USETMAXREG|FFMA2|FMUL2|FADD2|HMMA|REDG|UCGABAR|ACQBULK|MUFU\.TANH|MUFU\.EX2)")
per = collections.OrderedDict()
cur = None
for line in txt.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        per[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]+)", line)
    if m and cur is not None and KEEP.match(m.group(1)):
        per[cur][m.group(1)] += 1
tot = collections.Counter()
print("# Blackwell-specific SASS mnemonics per kernel of libodb200.so (cuobjdump -sass, sm_100a): tcgen05 MMAs (UTCHMMA, .2CTA = cta_group::2),")
print("# TMA loads / stores / reductions (UTMALDG / UTMASTG / UTMAREDG, UBLKCP = cp.async.bulk 1-D), tensor-memory ld/st (LDTM / STTM),")
print("# tcgen05.commit (UTCBAR, .MULTICAST), setmaxnreg (USETMAXREG), NVLS multimem (LDGMC = multimem.ld_reduce; multimem.st lowers to an ordinary STG on the multicast address),")
print("# packed fp32x2 math (FFMA2 / FMUL2 / FADD2).  Legacy tensor-core path would show as HMMA: none.\n")
for k, c in per.items():
    if not c:
        continue
    print(k)
    print("    " + "  ".join(f"{m} x{n}" for m, n in sorted(c.items())))
    for m, n in c.items():
        tot[re.sub(r"\..*", "", m) + (".2CTA" if ".2CTA" in m and m.startswith("UTCHMMA") else "")] += n
print("\n# totals: " + "  ".join(f"{m} x{n}" for m, n in sorted(tot.items())))
