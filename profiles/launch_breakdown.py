"""Sums an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel.  usage: python profiles/launch_breakdown.py launches.csv"""
import csv, sys, collections, re
rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot, cnt = collections.Counter(), collections.Counter()
for r in rows[1:]:
    v = float(r[vi].replace(",", ""))
    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[ui], 1.0) if r[ui] in ("ns", "us", "ms", "s") else 1e-3
    name = re.sub(r"\(.*", "", r[ki])[:60]
    tot[name] += v; cnt[name] += 1
for k, v in tot.most_common():
    print(f"{k:62s} launches {cnt[k]:5d}  total {v:10.1f} us  mean {v / cnt[k]:8.2f} us")
print(f"{'ALL':62s} launches {sum(cnt.values()):5d}  total {sum(tot.values()):10.1f} us")
