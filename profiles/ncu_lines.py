#!/usr/bin/env python3
"""Summarise an `ncu --page source --csv --print-source cuda,sass` dump: executed instructions and stall
samples per CUDA source line.  usage: ncu_lines.py dump.csv [top_n]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur_file = ""
data = []
tot_inst = tot_smp = 0
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if len(r) < 8 or r[0] in ("Line No", ""):
        continue
    try:
        line = int(r[0]); inst = int(r[7]); smp = int(r[6])
    except ValueError:
        continue
    data.append((inst, smp, cur_file, line, r[1].strip()[:100]))
    tot_inst += inst; tot_smp += smp
print(f"total warp-instructions {tot_inst}  samples {tot_smp}")
data.sort(reverse=True)
for inst, smp, f, line, src in data[:top]:
    print(f"{100.0 * inst / tot_inst:5.1f}% inst {100.0 * smp / max(1, tot_smp):5.1f}% smp  {f}:{line}  {src}")
