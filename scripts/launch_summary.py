"""Per-kernel summary (count, mean µs, share) of an `ncu --metrics gpu__time_duration.sum --csv` launch list."""
import collections
import csv
import sys

for path in sys.argv[1:]:
    with open(path) as fh:
        lines = [line for line in fh if line.startswith('"')]
    per = collections.defaultdict(list)
    for row in csv.DictReader(lines):
        value = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        value = value / 1000 if unit in ("ns", "nsecond") else (value * 1000 if unit in ("ms", "msecond") else value)
        per[row["Kernel Name"].split("(")[0]].append(value)
    total = sum(sum(v) for v in per.values())
    print(path)
    for name, v in sorted(per.items(), key=lambda kv: -sum(kv[1])):
        print(f"  {name[:60]:60s} n={len(v):4d} mean={sum(v) / len(v):8.2f} us  min={min(v):8.2f}  share={sum(v) / total:.3f}")
