#!/usr/bin/env python
"""Top stalled SASS instructions of a kernel from `ncu -i rep --page source --csv`.
usage: python profiles/hot_sass.py <rep.ncu-rep> <kernel-regex> [n]"""
import csv, subprocess, sys
rep, pat = sys.argv[1], sys.argv[2]
n = int(sys.argv[3]) if len(sys.argv) > 3 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + pat],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
H = rows[hi]
data = []
for r in rows[hi + 1:]:
    if len(r) != len(H) or r[0] == "Address":
        if r and r[0] == "Kernel Name" and data:
            break          # first launch only
        continue
    data.append(r)
ci = {h: i for i, h in enumerate(H)}
def I(r, h):
    try: return int(float(r[ci[h]] or 0))
    except ValueError: return 0
tot = sum(I(r, "# Samples") for r in data)
print("samples", tot, "sass instr", len(data), "warp-inst executed", sum(I(r, "Instructions Executed") for r in data))
stall = [h for h in H if h.startswith("stall_") and "Not Issued" not in h]
agg = {h[6:]: sum(I(r, h) for r in data) for h in stall}
print("stall totals:", {k: v for k, v in sorted(agg.items(), key=lambda x: -x[1]) if v})
for r in sorted(data, key=lambda r: -I(r, "# Samples"))[:n]:
    st = {h[6:]: I(r, h) for h in stall if I(r, h) > 0}
    print("%6d %7d  %-95s %s" % (I(r, "# Samples"), I(r, "Instructions Executed"), r[ci["Source"]][:95],
                                 dict(sorted(st.items(), key=lambda x: -x[1])[:3])))
