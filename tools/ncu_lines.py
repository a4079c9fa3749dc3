#!/usr/bin/env python
"""Summarise an ncu report per CUDA source line: stall reasons and hot lines.
usage: python tools/ncu_lines.py gpurun_out/prof.ncu-rep [top_n]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
hdr = rows[hi]
ix = {n: i for i, n in enumerate(hdr)}
agg = {}
cur_file = ""
for r in rows:
    if r and r[0] == "File Path":
        cur_file = r[1].rsplit("/", 1)[-1]
        continue
    if r and r[0].isdigit():
        try:
            key = (cur_file, int(r[0]))
            a = agg.setdefault(key, [r[1][:110], 0, 0, {}])
            a[1] += int(r[ix["# Samples"]]); a[2] += int(r[ix["Instructions Executed"]])
            for n in hdr:
                if n.startswith("stall_") and "Not Issued" not in n:
                    a[3][n] = a[3].get(n, 0) + int(r[ix[n]] or 0)
        except (ValueError, IndexError):
            pass
tot = sum(a[1] for a in agg.values()); toti = sum(a[2] for a in agg.values())
st = {}
for a in agg.values():
    for k, v in a[3].items():
        st[k] = st.get(k, 0) + v
print("samples", tot, "warp instructions", toti)
print("stalls:", ", ".join(f"{k[6:]} {100 * v / tot:.1f}%" for k, v in sorted(st.items(), key=lambda x: -x[1])[:9]))
for ln, a in sorted(agg.items(), key=lambda x: -x[1][1])[:top]:
    main = max(a[3].items(), key=lambda x: x[1])[0][6:] if a[3] else ""
    print(f"{ln[0][:14]:14s}:{ln[1]:5d} {100 * a[1] / tot:5.1f}% smp {100 * a[2] / toti:5.1f}% inst  [{main:14s}] {a[0]}")
