#!/usr/bin/env python3
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, mean us, share."""
import collections
import csv
import sys


def main(path, steps=8):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        v = v / 1000.0 if unit in ("ns", "nsecond") else v * 1000.0 if unit in ("ms", "msecond") else v
        agg.setdefault(row["Kernel Name"], []).append(v)
    total = sum(sum(v) for v in agg.values())
    print("%-72s %5s %9s %9s %6s" % ("kernel", "n", "mean_us", "us/step", "share"))
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print("%-72s %5d %9.1f %9.1f %5.1f%%" % (k[:72], len(v), sum(v) / len(v), sum(v) / steps, 100 * sum(v) / total))
    print("total us/step: %.1f" % (total / steps))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 8)
