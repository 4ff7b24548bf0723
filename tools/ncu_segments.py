#!/usr/bin/env python
"""Summarise an ncu report's SASS page for one kernel: runs of instructions with the same execution count, with their
dynamic instruction totals, stall samples and shared-memory wavefronts (actual vs ideal).
usage: ncu_segments.py report.ncu-rep [kernel-index] [--list FILE]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
ki = int(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2].isdigit() else 0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
ks, cur = [], None
for r in csv.reader(out.splitlines()):
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}
        ks.append(cur)
    elif r and r[0] == "Address":
        cur["hdr"] = r
    elif cur is not None and r:
        cur["rows"].append(r)
k = ks[ki]
h = k["hdr"]
iS, iE, iSrc = h.index("# Samples"), h.index("Instructions Executed"), h.index("Source")
iW, iWi = h.index("L1 Wavefronts Shared"), h.index("L1 Wavefronts Shared Ideal")
print(k["name"][:120])
print("samples", sum(int(r[iS]) for r in k["rows"]), "warp-instructions", sum(int(r[iE]) for r in k["rows"]))
if "--list" in sys.argv:
    with open(sys.argv[sys.argv.index("--list") + 1], "w") as f:
        for n, r in enumerate(k["rows"]):
            f.write(f"{n:5d} {int(r[iS]):6d} {int(r[iE]):9d} {r[iW]:>9} {r[iWi]:>9}  {r[iSrc].strip()}\n")
seg, cur = [], None
for n, r in enumerate(k["rows"]):
    e = int(r[iE])
    if cur and e == cur["e"]:
        cur["n"] += 1; cur["s"] += int(r[iS]); cur["end"] = n; cur["w"] += int(r[iW]); cur["wi"] += int(r[iWi])
    else:
        cur = {"start": n, "end": n, "e": e, "n": 1, "s": int(r[iS]), "w": int(r[iW]), "wi": int(r[iWi])}
        seg.append(cur)
for s in seg:
    if s["e"] > 0 and (s["s"] > 100 or s["n"] > 10) and s["e"] * s["n"] > 2e5:
        print(f"{s['start']:5d}-{s['end']:5d} n={s['n']:4d} exec={s['e']:9d} dyn={s['e'] * s['n'] / 1e6:7.2f}M samples={s['s']:6d} smem wavefronts={s['w'] / 1e6:.2f}M ideal={s['wi'] / 1e6:.2f}M")
