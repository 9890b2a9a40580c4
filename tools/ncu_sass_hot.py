"""Hot SASS instructions (stall samples) of an `ncu --page source --csv` export between two anchors.
Usage: ncu_sass_hot.py file.csv [min_samples] [start_row end_row]"""
import csv
import sys
rows = list(csv.reader(open(sys.argv[1], newline='')))
hdr = rows[1]; ix = {k: i for i, k in enumerate(hdr)}
data = rows[2:]
mins = int(sys.argv[2]) if len(sys.argv) > 2 else 120
lg = [i for i, r in enumerate(data) if 'MUFU.LG2' in r[1]]
print('MUFU.LG2 rows', lg)
start, end = (int(sys.argv[3]), int(sys.argv[4])) if len(sys.argv) > 4 else (lg[2] - 900, lg[3] + 1100)
stall_cols = [k for k in hdr if k.startswith('stall_') and 'Not Issued' not in k]
tot = sum(int(r[ix['# Samples']]) for r in data if r[ix['# Samples']].isdigit())
acc = 0
for i in range(max(0, start), min(end, len(data))):
    r = data[i]
    s = int(r[ix['# Samples']]) if r[ix['# Samples']].isdigit() else 0
    acc += s
    if s >= mins:
        st = sorted(((k[6:], int(r[ix[k]])) for k in stall_cols if r[ix[k]].isdigit() and int(r[ix[k]]) > 0), key=lambda kv: -kv[1])[:3]
        print(i, f"{100*s/tot:4.1f}%", r[1].strip()[:70], st, 'exec', r[ix['Instructions Executed']])
print('region share %.1f%%' % (100 * acc / tot), 'total samples', tot)
