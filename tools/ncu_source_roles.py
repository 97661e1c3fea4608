"""Summarises `ncu -i X.ncu-rep --page source --csv` for the warp-specialised conv kernel (dev / evidence tool): per launch, the
warp-stall samples of the three roles (producer = the code around UTMALDG, MMA issuer = around UTCHMMA, epilogue = around LDTM),
their executed warp-instructions, and the most-sampled instructions.
python tools/ncu_source_roles.py source_page.csv > profiles/r2_conv_ncu_source_roles.txt"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
kernels, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "rows": []}
        kernels.append(cur)
    elif cur is not None:
        if cur["hdr"] is None:
            cur["hdr"] = r
        else:
            cur["rows"].append(r)
seen = set()
for kn, k in enumerate(kernels):
    h, R = k["hdr"], k["rows"]
    isrc, isamp, iex = h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
    stall = [c for c in h if c.startswith("stall_") and "Not Issued" not in c]
    sig = tuple(r[isamp] for r in R)
    if sig in seen:      # the page lists every launch twice (two views)
        continue
    seen.add(sig)

    def idx(tag):
        return [i for i, r in enumerate(R) if tag in r[isrc] and int(r[iex] or 0) > 0]

    def agg(lo, hi):
        c, n, ex = collections.Counter(), 0, 0
        for r in R[max(lo, 0):hi]:
            for s in stall:
                v = int(r[h.index(s)] or 0)
                if v:
                    c[s[6:]] += v
            n += int(r[isamp] or 0)
            ex += int(r[iex] or 0)
        return n, ex, dict(c.most_common(6))

    ld, mma, tm = idx("UTMALDG"), idx("UTCHMMA"), idx("LDTM")
    total = sum(int(r[isamp] or 0) for r in R)
    mma_n = sum(int(R[i][iex]) for i in mma)
    print(f"=== launch {kn // 2}: {total} samples, {mma_n} UTCHMMA, {sum(int(R[i][iex]) for i in tm)} LDTM warp-instructions")
    if ld and mma and tm:
        p0, p1 = ld[0] - 120, ld[-1] + 80
        m0, m1 = mma[0] - 650, mma[-1] + 90
        for name, (a, b) in (("producer", (p0, p1)), ("epilogue", (p1, m0)), ("MMA issuer", (m0, m1)), ("exit / shared wait loop", (m1, len(R)))):
            n, ex, c = agg(a, b)
            print(f"  {name:24s} samples {n:5d}  warp-instructions {ex:9d}  stalls {c}")
    top = sorted(R, key=lambda r: -int(r[isamp] or 0))[:12]
    for r in top:
        s = {c[6:]: int(r[h.index(c)] or 0) for c in stall}
        s = {a: b for a, b in s.items() if b}
        print(f"    {int(r[isamp]):5d}  x{int(r[iex] or 0):<8d} {r[isrc][:70]:70s} {s}")
