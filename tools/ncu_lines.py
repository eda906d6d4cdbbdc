#!/usr/bin/env python
"""Development tool: aggregate an `ncu --page source --csv` dump by source line: share of stall samples, of executed
instructions and shared-memory wavefronts (actual vs ideal) for the hottest lines of each profiled kernel."""
import csv, sys
path, kids = sys.argv[1], [int(x) for x in sys.argv[2].split(",")]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
rows = list(csv.reader(open(path)))
kernels, cur = [], None
for r in rows:
    if r and r[0] == "Function Name":
        cur = {"name": r[1], "lines": []}; kernels.append(cur)
    elif r and r[0] == "Line No":
        cur["hdr"] = r
    elif cur is not None and "hdr" in cur and r and r[0] not in ("File Path", "") and r[0].isdigit():
        cur["lines"].append(r)
for kid in kids:
    k = kernels[kid]; h = k["hdr"]
    iS, iI, iW, iWi = h.index("# Samples"), h.index("Instructions Executed"), h.index("L1 Wavefronts Shared"), h.index("L1 Wavefronts Shared Ideal")
    tot_s = sum(int(r[iS]) for r in k["lines"]) or 1; tot_i = sum(int(r[iI]) for r in k["lines"]) or 1
    print("=====", kid, k["name"][:70], "samples", tot_s, "inst", tot_i)
    L = sorted(k["lines"], key=lambda r: -int(r[iS]))[:top]
    for r in sorted(L, key=lambda r: int(r[0])):
        print(f"{r[0]:>4} smp {100*int(r[iS])/tot_s:5.1f}% inst {100*int(r[iI])/tot_i:5.1f}% wf {r[iW]:>9} ideal {r[iWi]:>9} | {r[1].strip()[:100]}")
