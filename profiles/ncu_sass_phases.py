#!/usr/bin/env python3
"""Summarise an `ncu --page source --csv --print-source sass` dump of ONE kernel launch: executed warp
instructions and stall samples per program phase (phases = runs of SASS between BAR.SYNC instructions) and per
opcode.  usage: ncu_sass_phases.py dump.csv"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = None
kern = 0
phase = 0
ph_inst = collections.Counter(); ph_smp = collections.Counter(); ph_ops = collections.defaultdict(collections.Counter)
op_inst = collections.Counter(); op_smp = collections.Counter()
stall_cols = {}
ph_stall = collections.defaultdict(collections.Counter)
for r in rows:
    if r and r[0] == "Kernel Name":
        kern += 1
        continue
    if kern != 1:
        continue
    if r and r[0] == "Address":
        hdr = r
        continue
    if hdr is None or len(r) < 8:
        continue
    try:
        inst = int(r[5]); smp = int(r[4])
    except ValueError:
        continue
    sass = r[1].strip()
    toks = sass.split()
    op = toks[1] if toks[0].startswith("@") else toks[0]
    base = op.split(".")[0]
    if base == "BAR":
        phase += 1
    ph_inst[phase] += inst; ph_smp[phase] += smp; ph_ops[phase][base] += inst
    op_inst[base] += inst; op_smp[base] += smp
    for i in range(32, len(r)):
        try:
            v = int(r[i])
        except ValueError:
            continue
        if v:
            ph_stall[phase][hdr[i] if i < len(hdr) else str(i)] += v
ti = sum(ph_inst.values()); ts = sum(ph_smp.values())
print(f"total warp-instructions {ti}  stall samples {ts}")
for p in sorted(ph_inst):
    top = ", ".join(f"{o} {100*c/ph_inst[p]:.0f}%" for o, c in ph_ops[p].most_common(6)) if ph_inst[p] else ""
    st = ", ".join(f"{k.replace('stall_','')} {v}" for k, v in ph_stall[p].most_common(4))
    print(f"phase {p:2d}: {100*ph_inst[p]/ti:5.1f}% inst {100*ph_smp[p]/max(ts,1):5.1f}% smp | {top} | {st}")
print("opcodes:")
for o, c in op_inst.most_common(25):
    print(f"  {o:10s} {100*c/ti:5.1f}% inst {100*op_smp[o]/max(ts,1):5.1f}% smp")
