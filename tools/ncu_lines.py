"""Aggregate an `ncu --page source --csv --print-source cuda,sass` export per CUDA source line:
samples, executed warp instructions and the dominant stall reasons. Usage: ncu_lines.py file.csv [top_n]"""
import csv
import sys
from collections import defaultdict

rows = list(csv.reader(open(sys.argv[1], newline="")))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 60
cur_file, hdr = None, None
lines = []
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r and r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or len(r) < len(hdr) - 2 or r[0] == "":
        continue
    d = dict(zip(hdr, r))
    try:
        samples = int(d["# Samples"]); inst = int(d["Instructions Executed"])
    except ValueError:
        continue
    stalls = {k[6:]: int(v) for k, v in d.items() if k.startswith("stall_") and "Not Issued" not in k and v.isdigit() and int(v) > 0}
    lines.append((cur_file, int(r[0]), r[1].strip(), samples, inst, stalls))
tot_s = sum(l[3] for l in lines); tot_i = sum(l[4] for l in lines)
print(f"total samples {tot_s}  warp instructions {tot_i}")
for f, ln, src, s, i, st in sorted(lines, key=lambda l: -l[3])[:top]:
    stt = " ".join(f"{k}:{v}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:4])
    print(f"{100*s/tot_s:5.1f}% smp {100*i/tot_i:5.1f}% ins  {f}:{ln:<5d} {src[:70]:70s} | {stt}")
