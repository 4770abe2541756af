"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel."""
import collections
import csv
import sys

lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
agg = collections.OrderedDict()
for row in csv.DictReader(lines):
    agg.setdefault(row["Kernel Name"].split("(")[0], []).append(float(row["Metric Value"].replace(",", "")))
tot = sum(sum(v) for v in agg.values())
print("%-60s %5s %12s %12s %12s %7s" % ("kernel", "n", "mean_ns", "min_ns", "max_ns", "share"))
for k, v in agg.items():
    print("%-60s %5d %12.0f %12.0f %12.0f %6.1f%%" % (k[:60], len(v), sum(v) / len(v), min(v), max(v), 100 * sum(v) / tot))
