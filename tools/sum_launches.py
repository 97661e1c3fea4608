"""Aggregate an ncu launch list (--metrics gpu__time_duration.sum --csv) by kernel name: count, total, mean (debug aid)."""
import csv, sys, collections, re
rows = [r for r in csv.reader(open(sys.argv[1], errors="ignore")) if len(r) > 5]
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
H = rows[hdr]; kn = H.index("Kernel Name"); mv = H.index("Metric Value"); mu = H.index("Metric Unit")
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
agg = collections.OrderedDict()
n = 0
for r in rows[hdr + 1:]:
    n += 1
    if n <= skip:
        continue
    name = re.sub(r"\(.*", "", r[kn])[:70]
    v = float(r[mv].replace(",", ""))
    v = v / 1000.0 if r[mu] in ("ns", "nsecond") else v
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
for name, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    print(f"{t:10.1f} us {100 * t / tot:5.1f}%  x{c:<5d} mean {t / c:8.1f}  {name}")
print(f"total {tot:.1f} us over {sum(a[0] for a in agg.values())} launches")
