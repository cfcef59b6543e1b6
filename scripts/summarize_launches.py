"""Condenses an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel count / mean / share."""
import csv
import sys
from collections import defaultdict

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 14 and r[0].isdigit()]
agg = defaultdict(list)
for r in rows:
    name = r[4].split("(")[0].replace("void ", "")
    if "<" in name:
        name = name[: name.index("<")] + "<" + name[name.index("<") + 1:][:48]
    val = float(r[14].replace(",", ""))
    unit = r[13]
    us = val / 1000.0 if unit in ("ns", "nsecond") else (val if unit in ("us", "usecond") else val * 1000.0)
    agg[name].append(us)
total = sum(sum(v) for v in agg.values())
print(f"| kernel | launches | mean us | total us | share |\n|---|---:|---:|---:|---:|")
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print(f"| `{k}` | {len(v)} | {sum(v) / len(v):.1f} | {sum(v):.1f} | {sum(v) / total:.1%} |")
