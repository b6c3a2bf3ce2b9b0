#!/usr/bin/env python
"""One training iteration out of an ncu launch list: ordered kernels with grid and duration, and per-kernel totals."""
import collections
import csv
import sys


def main(path, detail=True):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    hdr = rows[hi]
    kn, mv, gs = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size")
    L = [(r[kn].split("(")[0].replace("void ", "").replace("glis::", ""), r[gs], float(r[mv].replace(",", "")) / 1e3)
         for r in rows[hi + 1:] if len(r) > mv]
    idx = [i for i, l in enumerate(L) if l[0].startswith("rmsprop")]
    s, e = idx[-3] + 1, idx[-1] + 1          # the last complete iteration (two optimizer steps)
    agg = collections.defaultdict(lambda: [0, 0.0])
    tot = 0.0
    for i in range(s, e):
        n, g, t = L[i]
        tot += t
        agg[n[:56]][0] += 1
        agg[n[:56]][1] += t
        if detail:
            print("%3d %-56s %-16s %7.1f" % (i - s, n[:56], g, t))
    print()
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-58s %3d %8.1f %5.1f%%" % (k, v[0], v[1], 100 * v[1] / tot))
    print("total %.1f us over %d launches" % (tot, e - s))


if __name__ == "__main__":
    main(sys.argv[1], len(sys.argv) < 3)
