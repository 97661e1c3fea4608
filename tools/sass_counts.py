"""Evidence helper: counts of the Blackwell-specific SASS mnemonics per kernel of liby5obb.so (cuobjdump -sass), written to
profiles/r2_sass_counts.txt.  python tools/sass_counts.py"""
import re
import subprocess
import sys
from collections import Counter, defaultdict
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
so = ROOT / "yolov5_obb_b200" / "liby5obb.so"
out = subprocess.run(["cuobjdump", "-sass", str(so)], capture_output=True, text=True).stdout
pat = {"UTCHMMA (tcgen05.mma kind::f16)": r"\bUTCHMMA\b", "UTCBAR (tcgen05.commit)": r"\bUTCBAR\b", "LDTM (tcgen05.ld)": r"\bLDTM\b",
       "UTMALDG (TMA load)": r"\bUTMALDG\b", "UTMASTG (TMA store)": r"\bUTMASTG\b", "UTMAREDG (TMA reduce-add store)": r"\bUTMAREDG\b", "UTCATOMSWS/alloc (tcgen05.alloc)": r"\bUTCATOMSWS\b",
       "SYNCS (mbarrier)": r"\bSYNCS\b", "ELECT": r"\bELECT\b", "BRA.U.ANY (per-lane serialisation loops)": r"BRA\.U\.ANY",
       "HMMA (legacy mma.sync)": r"\bHMMA\b", "ACQBULK/griddepcontrol": r"ACQBULK|\bDEPBAR\b"}
cur, per = None, defaultdict(Counter)
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        continue
    if cur is None:
        continue
    for name, rx in pat.items():
        if re.search(rx, line):
            per[cur][name] += 1
def demangle(s):
    d = subprocess.run(["c++filt", s], capture_output=True, text=True).stdout.strip()
    m = re.search(r"(\w+)\(", d.replace("(anonymous namespace)::", ""))
    return m.group(1) if m else d[:60]
lines = [f"SASS mnemonic counts per kernel of {so.name} (cuobjdump -sass, sm_100a); kernels without any of them are omitted", ""]
for fn, c in sorted(per.items(), key=lambda kv: -sum(kv[1].values())):
    if not any(k.startswith(("UTC", "LDTM", "UTMA")) for k in c):
        continue
    lines.append(demangle(fn))
    for name in pat:
        if c[name]:
            lines.append(f"    {name:45s} {c[name]}")
tot = Counter()
for c in per.values():
    tot.update(c)
lines += ["", "whole library"] + [f"    {name:45s} {tot[name]}" for name in pat]
(ROOT / "profiles" / "r2_sass_counts.txt").write_text("\n".join(lines) + "\n")
print("\n".join(lines))
