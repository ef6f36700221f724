"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: time share per kernel."""
import collections
import csv
import re
import sys


def main(path, skip=0, take=None):
    rows = []
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    rd = csv.reader(lines)
    hdr = next(rd)
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    for r in rd:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        u = r[ui]
        us = v / 1e3 if u in ("ns", "nsecond") else (v if u in ("us", "usecond") else v * 1e3)
        name = re.sub(r"\(.*", "", r[ki]).replace("void ", "").replace("<unnamed>::", "")
        rows.append((name, us))
    rows = rows[skip: skip + take if take else None]
    tot = sum(u for _, u in rows)
    agg = collections.defaultdict(lambda: [0, 0.0])
    for n, u in rows:
        agg[n][0] += 1
        agg[n][1] += u
    print(f"{len(rows)} launches, {tot:.1f} us total")
    for n, (c, u) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{100 * u / tot:5.1f}%  {u:10.1f} us  {c:5d} x  avg {u / c:8.1f} us  {n[:90]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0, int(sys.argv[3]) if len(sys.argv) > 3 else None)
