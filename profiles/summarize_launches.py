"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel count, mean us, share.
usage: python profiles/summarize_launches.py gpurun_out/launches.csv [skip_first_n] > profiles/xxx.md"""
import csv
import sys
from collections import OrderedDict

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
h = rows[0]
ki, vi = h.index("Kernel Name"), h.index("Metric Value")
agg = OrderedDict()
for r in rows[1 + skip:]:
    name = r[ki].split("(")[0].replace("<unnamed>::", "").replace("void ", "")[:80]
    d = agg.setdefault(name, [0, 0.0])
    d[0] += 1
    d[1] += float(r[vi].replace(",", "")) / 1e3
tot = sum(v[1] for v in agg.values())
print(f"| kernel | launches | mean us | total us | share |\n|---|---|---|---|---|")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{k}` | {n} | {t / n:.1f} | {t:.1f} | {100 * t / tot:.1f}% |")
print(f"\ntotal {tot:.1f} us over {sum(v[0] for v in agg.values())} launches (ncu-serialised, cold-cache: compare shares)")
