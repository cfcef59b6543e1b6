"""Summarises an `ncu --page source --csv` dump: stall-reason totals, instruction mix and the hottest SASS lines."""
import csv
import re
import sys
from collections import Counter

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
rows = list(csv.reader(open(path)))
hi = next(i for i, r in enumerate(rows) if "Source" in r and "# Samples" in r)
hdr = rows[hi]
col = {k: i for i, k in enumerate(hdr)}
body = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
stalls = [k for k in hdr if k.startswith("stall_") and "Not Issued" not in k]
tot = Counter()
mix = Counter()
inst_total = 0
for r in body:
    for k in stalls:
        try:
            tot[k] += int(r[col[k]])
        except ValueError:
            pass
    try:
        n = int(r[col["Instructions Executed"]])
    except ValueError:
        n = 0
    inst_total += n
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[col["Source"]])
    if m:
        mix[m.group(2).split(".")[0]] += n
samples = sum(tot.values())
print(f"SASS lines: {len(body)}, warp instructions executed: {inst_total}, stall samples: {samples}")
print("stall reasons:", ", ".join(f"{k[6:]}={v / max(samples, 1):.1%}" for k, v in tot.most_common(8)))
print("instruction mix:", ", ".join(f"{k}={v / max(inst_total, 1):.1%}" for k, v in mix.most_common(14)))
print(f"hottest {top} SASS lines by samples:")
key = lambda r: -int(r[col["# Samples"]] or 0)  # noqa: E731
for r in sorted(body, key=key)[:top]:
    rs = {k[6:]: int(r[col[k]] or 0) for k in stalls}
    main = max(rs, key=rs.get)
    print(f"  {r[col['# Samples']]:>6s}  {main:14s} {r[col['Source']][:90]}")
