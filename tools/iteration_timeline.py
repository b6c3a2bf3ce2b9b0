#!/usr/bin/env python
"""One iteration of an `ncu --metrics gpu__time_duration.sum --csv` launch list, in launch order, per stream.
usage: python tools/iteration_timeline.py launches.csv [iteration index from the end, default 1]"""
import csv
import sys


def main(path, back=1):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    hdr = rows[hi]
    kn, mv, mu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    st = hdr.index("Stream") if "Stream" in hdr else None
    gs, bs = hdr.index("Grid Size"), hdr.index("Block Size")
    ev = []
    for r in rows[hi + 1:]:
        if len(r) <= mv:
            continue
        v = float(r[mv].replace(",", ""))
        v = v / 1e3 if r[mu] == "ns" else (v * 1e3 if r[mu] == "ms" else v)
        ev.append((r[kn].split("(")[0].replace("glis::", "").replace("void ", ""), r[st] if st is not None else "0", v, r[gs], r[bs]))
    # iteration boundaries: the rmsprop kernel runs twice per iteration (D, then G)
    opt = [i for i, e in enumerate(ev) if "rmsprop" in e[0]]
    ends = opt[1::2]
    e1 = ends[-back]
    e0 = ends[-back - 1]
    it = ev[e0 + 1:e1 + 1]
    streams = {}
    for e in it:
        streams.setdefault(e[1], [0, 0.0])
        streams[e[1]][0] += 1
        streams[e[1]][1] += e[2]
    print("iteration: %d launches, %.1f us serialised; per stream: %s" % (
        len(it), sum(e[2] for e in it), {k: (v[0], round(v[1], 1)) for k, v in streams.items()}))
    for i, e in enumerate(it):
        print("%3d  s%-3s %8.1f us  %-48s grid %-18s block %s" % (i, e[1], e[2], e[0][:48], e[3], e[4]))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 1)
