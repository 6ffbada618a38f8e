"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel: count, total, average, share."""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1], errors="replace")))
hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
h = rows[hi]
ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) <= vi or not r[vi].replace(",", "").replace(".", "").isdigit():
        continue
    name = re.sub(r"\(.*$", "", r[ki])
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\(anonymous namespace\)::", "", name)[:90]
    v = float(r[vi].replace(",", "")) / (1e3 if r[ui] in ("ns", "nsecond") else 1.0)
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1; a[1] += v
tot = sum(v[1] for v in agg.values())
print("%-92s %6s %11s %9s %7s" % ("kernel", "count", "total us", "avg us", "share"))
for n, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print("%-92s %6d %11.1f %9.2f %6.1f%%" % (n, c, t, t / c, 100 * t / tot))
print("total %.1f us over %d launches" % (tot, sum(v[0] for v in agg.values())))
