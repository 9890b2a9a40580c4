"""Aggregate an `ncu --page source --csv --print-source cuda,sass` export of k_subdomain per kernel region
(line ranges of jj_subdomain.cu given as name:lo-hi arguments). Usage: ncu_regions.py file.csv blocks steps name:lo-hi ..."""
import csv
import sys
rows = list(csv.reader(open(sys.argv[1], newline='')))
blocks, steps = int(sys.argv[2]), int(sys.argv[3])
cur = None; hdr = None; agg = {}
tot_s = tot_i = 0
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur = r[1].split('/')[-1]; continue
    if r and r[0] == "Line No":
        hdr = r; continue
    if hdr is None or len(r) < len(hdr) - 2 or r[0] == "":
        continue
    d = dict(zip(hdr, r))
    try:
        s = int(d["# Samples"]); i = int(d["Instructions Executed"])
    except ValueError:
        continue
    st = {k[6:]: int(v) for k, v in d.items() if k.startswith("stall_") and "Not Issued" not in k and v.isdigit()}
    agg[(cur, int(r[0]))] = (s, i, st)
    tot_s += s; tot_i += i
f = 'jj_subdomain.cu'
def region(name, pred):
    S = I = 0; ST = {}
    for k, (s, i, st) in agg.items():
        if pred(k):
            S += s; I += i
            for a, b in st.items():
                ST[a] = ST.get(a, 0) + b
    top = sorted(ST.items(), key=lambda kv: -kv[1])[:7]
    print(f"{name:24s} samples {100*S/tot_s:5.1f}%  instr {100*I/tot_i:5.1f}% ({I/blocks/steps/1000:6.1f}k per block-step)  " +
          " ".join(f"{a}:{100*b/max(S,1):.0f}%" for a, b in top if a != 'branch_resolving' or b < S))
print('total samples', tot_s, 'warp instructions per block-step %.1fk' % (tot_i / blocks / steps / 1000))
region('jj_device.cuh', lambda k: k[0] == 'jj_device.cuh')
for spec in sys.argv[4:]:
    name, rng = spec.split(':'); lo, hi = map(int, rng.split('-'))
    region(name, lambda k: k[0] == f and lo <= k[1] <= hi)
region('other files', lambda k: k[0] not in (f, 'jj_device.cuh'))
