"""Print the hottest SASS instructions (warp-stall samples) from `ncu --page source --csv` output."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = next(r for r in rows if "# Samples" in r)
data = [r for r in rows if len(r) == len(hdr) and r[hdr.index("# Samples")].isdigit()]
si, src = hdr.index("# Samples"), hdr.index("Source")
tot = sum(int(r[si]) for r in data) or 1
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
print("total samples", tot, "instructions", len(data))
for r in sorted(data, key=lambda r: -int(r[si]))[: int(sys.argv[2]) if len(sys.argv) > 2 else 25]:
    st = sorted(((int(r[i]), hdr[i]) for i in stall_cols), reverse=True)[:2]
    print(f"{int(r[si]):7d} {100 * int(r[si]) / tot:5.1f}%  {r[src].strip()[:72]:72s} {st}")
