#!/usr/bin/env python
"""Top stalled SASS instructions of `ncu -i X.ncu-rep --page source --csv` (first kernel, or --kernel N)."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = next(r for r in rows if "Source" in r and "# Samples" in r)
start = rows.index(hdr) + 1
i_src, i_s, i_ex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
data = []
for r in rows[start:]:
    if len(r) != len(hdr):
        break
    data.append(r)
tot = sum(int(r[i_s]) for r in data)
print("total samples", tot, "instructions", len(data))
top = sorted(enumerate(data), key=lambda t: -int(t[1][i_s]))[:n]
for idx, r in sorted(top):
    st = {hdr[i][6:]: int(r[i]) for i in stall_cols if int(r[i]) > 0}
    st = dict(sorted(st.items(), key=lambda kv: -kv[1])[:3])
    print(f"{idx:5d} {100 * int(r[i_s]) / tot:5.1f}% x{r[i_ex]:>8s}  {r[i_src].strip()[:72]:72s} {st}")
