"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total and
share of the captured window (cold-cache, serialised launches: compare SHARES, not absolutes)."""
import csv
import re
import sys
from collections import defaultdict

rows = []
with open(sys.argv[1], newline="") as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    ns = v * {"ns": 1, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6, "nsecond": 1, "second": 1e9}.get(unit, 1)
    name = re.sub(r"\(.*", "", r["Kernel Name"])
    name = re.sub(r"<unnamed>::|\(anonymous namespace\)::|void ", "", name)
    rows.append((name, ns))
# optional window: keep the launches of train steps [first, last) counted by step_tick_kernel (one per step)
if len(sys.argv) >= 4:
    first, last = int(sys.argv[2]), int(sys.argv[3])
    ticks = [i for i, (n, _) in enumerate(rows) if "step_tick_kernel" in n]
    lo = ticks[first] if first < len(ticks) else 0
    hi = ticks[last] if last < len(ticks) else len(rows)
    rows = rows[lo:hi]
    print(f"window: train steps {first}..{last - 1} ({last - first} steps)")
tot = sum(ns for _, ns in rows)
agg = defaultdict(lambda: [0, 0.0])
for n, ns in rows:
    agg[n][0] += 1
    agg[n][1] += ns
print(f"launches {len(rows)}  total {tot/1e3:.1f} us")
print(f"{'kernel':70s} {'count':>6s} {'total_us':>10s} {'avg_us':>9s} {'share':>7s}")
for n, (c, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{n[:70]:70s} {c:6d} {ns/1e3:10.1f} {ns/1e3/c:9.2f} {100*ns/tot:6.1f}%")
