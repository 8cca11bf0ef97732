#!/usr/bin/env python3
"""cuobjdump -sass xevd_b200/libxevd_b200.so | python profiles/sass_evidence.py > profiles/rN/sass_evidence.txt
Per-kernel counts of the SASS mnemonics that show what the data movement and the arithmetic are made of."""
import collections
import re
import sys

txt = sys.stdin.read()
funcs = re.split(r"\n\s*Function : ", txt)
print("# SASS evidence for xevd_b200/libxevd_b200.so (cuobjdump -sass, nvcc 12.9, sm_100a) - instruction counts per kernel of the mnemonics that show")
print("# what the data movement and the arithmetic are made of: UTMALDG = TMA tensor load (cp.async.bulk.tensor), UBLKCP = bulk copy (cp.async.bulk),")
print("# SYNCS = mbarrier, IDP.2A = packed s16x2 . s8x2 dot product, VIADDMNMX / VIMNMX = DPX min/max-add (packed s16x2 clip), I2IP = saturating pack.")
print("# regenerate: cuobjdump -sass xevd_b200/libxevd_b200.so | python profiles/sass_evidence.py\n")
tot = collections.Counter()
KEYS = ("UTMALDG", "UBLKCP", "SYNCS", "IDP.2A", "IDP.4A", "VIADDMNMX", "VIMNMX", "I2IP", "IMAD", "LDS", "STS", "BAR")
for f in funcs[1:]:
    name = f.split("\n", 1)[0].strip()
    c = collections.Counter()
    for m in re.finditer(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", f, re.M):
        op = m.group(1)
        for key in KEYS:
            if op.startswith(key):
                c[key] += 1
        c["total"] += 1
    tot.update(c)
    short = re.sub(r"^_ZN2xb", "", name)[:70]
    print(f"{short:72s} total {c['total']:6d}  UTMALDG {c['UTMALDG']:3d}  UBLKCP {c['UBLKCP']:2d}  SYNCS {c['SYNCS']:3d}  IDP.2A {c['IDP.2A']:5d}  VIADDMNMX {c['VIADDMNMX']:4d}  "
          f"VIMNMX {c['VIMNMX']:4d}  I2IP {c['I2IP']:4d}  IMAD {c['IMAD']:5d}")
print("\nlibrary total: " + "  ".join(f"{k} {v}" for k, v in sorted(tot.items())))
