#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel totals for ONE bench step.
usage: tools/launch_summary.py launches.csv [step_index]   (a period runs from one match_init_kernel launch to the next = one step worth of launches in steady state)"""
import collections
import csv
import re
import sys


def main():
    path = sys.argv[1]
    step = int(sys.argv[2]) if len(sys.argv) > 2 else -2
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    names = [r["Kernel Name"] for r in rows]
    starts = [i for i, n in enumerate(names) if "match_init_kernel" in n]   # one per step; with prefetching the scan registration of a later step interleaves, every period still holds one launch set
    a = starts[step]; b = starts[step + 1] if step + 1 < len(starts) and step + 1 != 0 else len(rows)
    agg = collections.OrderedDict(); tot = 0.0
    for r in rows[a:b]:
        n = re.sub(r"\(.*", "", r["Kernel Name"]); n = re.sub(r"<.*", "", n)
        d = float(r["Metric Value"])
        agg.setdefault(n, [0.0, 0]); agg[n][0] += d; agg[n][1] += 1; tot += d
    print("step %d of %d: %d launches, %.1f us of kernel time (ncu: serialised, cold caches -> compare SHARES)" % (step, len(starts), b - a, tot / 1e3))
    for n, (d, c) in sorted(agg.items(), key=lambda x: -x[1][0]):
        print("%-52s %9.1f us  x%3d  %5.1f%%" % (n[:52], d / 1e3, c, 100 * d / tot))


if __name__ == "__main__":
    main()
