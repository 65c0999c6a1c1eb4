"""Summarise `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass` per CUDA source line:
stall samples, executed warp instructions and the dominant stall reason.  Usage: ncu_lines.py file.csv [top]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
cols = rows[hdr]
ci = {c: i for i, c in enumerate(cols)}
i_s, i_n, i_x = cols.index("# Samples"), cols.index("Warp Stall Sampling (Not-issued Samples)"), cols.index("Instructions Executed")
stall_cols = [i for i, c in enumerate(cols) if c.startswith("stall_") and "Not Issued" not in c]
lines = []
for r in rows[hdr + 1:]:
    if len(r) < len(cols) or not r[0].strip().isdigit():
        continue
    try:
        smp, inst = int(r[i_s]), int(r[i_x])
    except ValueError:
        continue
    st = sorted(((int(r[i]) if r[i].isdigit() else 0, cols[i]) for i in stall_cols), reverse=True)[:2]
    lines.append((smp, inst, int(r[0]), r[1].strip()[:110], st))
tot_s, tot_i = sum(l[0] for l in lines) or 1, sum(l[1] for l in lines) or 1
print(f"total samples {tot_s}  warp-instructions {tot_i}")
for smp, inst, ln, src, st in sorted(lines, reverse=True)[:top]:
    print(f"{100*smp/tot_s:5.1f}% smp {100*inst/tot_i:5.1f}% inst  L{ln:<4d} {src}   [{', '.join(f'{n}:{c}' for c, n in st if c)}]")
