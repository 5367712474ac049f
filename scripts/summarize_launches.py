"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: time per kernel family over the LAST step
(launches after the final occurrence of the frontend kernel), shares, counts."""
import csv
import re
import sys
from collections import OrderedDict


def main(path):
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if l.startswith('"')]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        ns = v * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1)
        rows.append((r["Kernel Name"], ns))
    if not rows:
        print("no rows")
        return
    last = max(i for i, (k, _) in enumerate(rows) if "frontend_kernel" in k)
    step = rows[last:]
    fam = OrderedDict()
    for k, ns in step:
        name = re.sub(r"<.*", "", k.split("(")[0]).strip()
        if "gemm_f16_kernel" in k:
            m = re.search(r"gemm_f16_kernel<\(int\)(\d), \(int\)(\d+), \(int\)(\d+), \(int\)(\d+), \(int\)(\d+)>", k)
            name = "gemm_f16_kernel" + (f"<cg{m.group(1)},n{m.group(2)},epi{m.group(5)}>" if m else "")
        a = fam.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += ns
    tot = sum(v[1] for v in fam.values())
    print(f"last step: {len(step)} launches, {tot / 1e6:.3f} ms (serialised, cold-cache ncu times)")
    for name, (n, ns) in sorted(fam.items(), key=lambda kv: -kv[1][1]):
        print(f"{ns / 1e6:9.3f} ms  {100 * ns / tot:5.1f} %  x{n:<4d} {name}")


if __name__ == "__main__":
    main(sys.argv[1])
