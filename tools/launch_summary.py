"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel (not part of the product)."""
import collections, csv, sys
path = sys.argv[1]
lines = [l for l in open(path) if not l.startswith("==")]
r = csv.reader(lines)
hdr = next(r)
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
seq = []
for row in r:
    if len(row) <= vi:
        continue
    v = float(row[vi].replace(",", ""))
    u = row[ui]
    v = v / 1000 if u == "ns" else (v * 1000 if u == "ms" else v)
    seq.append((row[ki].split("(")[0][:64], v))
tot = sum(v for _, v in seq)
agg = collections.OrderedDict()
for n, v in seq:
    a = agg.setdefault(n, [0, 0.0]); a[0] += 1; a[1] += v
print(f"total {tot:.1f} us over {len(seq)} launches")
print("| share | total us | launches | avg us | kernel |\n|---:|---:|---:|---:|---|")
for n, (c, v) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"| {v/tot*100:.2f}% | {v:.1f} | {c} | {v/c:.2f} | `{n}` |")
if len(sys.argv) > 2:
    for n, v in seq[int(sys.argv[2]):int(sys.argv[3])]:
        print(f"{v:9.2f}  {n}")
