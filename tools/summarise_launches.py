"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count / total / average / share.
usage: python tools/summarise_launches.py launches.csv"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
h = rows[hi]
kn, mv, mu = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
agg = collections.OrderedDict()
n = 0
for r in rows[hi + 1:]:
    if len(r) <= mv:
        continue
    v = float(r[mv].replace(",", ""))
    v = v / 1000 if r[mu] == "ns" else v * 1000 if r[mu] == "ms" else v
    name = r[kn].split("(")[0]
    a = agg.setdefault(name, [0, 0.0, 0.0])
    a[0] += 1
    a[1] += v
    a[2] = max(a[2], v)
    n += 1
setup = {"k_gens_tables", "k_build_comb", "k_point_tables", "k_mimc_sponge"}
tot = sum(a[1] for k, a in agg.items() if k not in setup)
print("launches %d  non-setup device time %.0f us (cold-cache, serialised: compare shares)" % (n, tot))
for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
    print("%-28s n=%4d sum=%9.1f avg=%8.1f max=%8.1f share=%5.1f%%" % (k, a[0], a[1], a[1] / a[0], a[2], 100 * a[1] / tot if k not in setup else 0))
