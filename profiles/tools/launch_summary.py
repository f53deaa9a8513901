"""Groups an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel.  python profiles/tools/launch_summary.py launches.csv"""
import collections
import csv
import re
import sys

rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr = rows[0]
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[1:]:
    name = re.sub(r"\(.*", "", r[ik]).replace("void ", "").replace("nimmt::", "").replace("at::", "")[:60]
    v = float(r[iv].replace(",", ""))
    v = v / 1000.0 if r[iu] in ("ns", "nsecond") else v          # -> us
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{name:62s} n={n:5d} avg={t / n:10.2f} us  share={100 * t / tot:5.1f}%")
