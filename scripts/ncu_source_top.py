"""Top warp-stall lines of one kernel from an ncu report:  ncu -i X.ncu-rep --page source --csv | python scripts/ncu_source_top.py [kernel-index] [N]"""
import csv
import sys

which = int(sys.argv[1]) if len(sys.argv) > 1 else 0
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 30
rows = list(csv.reader(sys.stdin))
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
starts.append(len(rows))
s, e = starts[which], starts[which + 1]
print(rows[s][1][:110])
h = rows[s + 1]
i_src, i_s = h.index("Source"), h.index("Warp Stall Sampling (All Samples)")
stalls = [k for k in h if k.startswith("stall_") and "Not Issued" not in k]
data = []
for r in rows[s + 2:e]:
    if len(r) < len(h):
        continue
    data.append((int(r[i_s]), r[i_src].strip(), {k: int(r[h.index(k)]) for k in stalls}))
tot = sum(d[0] for d in data)
agg = {}
for n, src, st in data:
    for k, v in st.items():
        agg[k] = agg.get(k, 0) + v
print("total samples", tot, sorted(agg.items(), key=lambda x: -x[1])[:8])
for n, src, st in sorted(data, key=lambda x: -x[0])[:topn]:
    top = sorted(st.items(), key=lambda x: -x[1])[:2]
    print(f"{n:6d} {100 * n / tot:5.1f}%  {src[:72]:72s} {top}")
