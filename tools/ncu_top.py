#!/usr/bin/env python
"""Top stalled SASS instructions of each kernel in an ncu report (needs -lineinfo / --import-source on).
  python tools/ncu_top.py gpurun_out/prof.ncu-rep [N]
"""
import csv, io, subprocess, sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                     capture_output=True, text=True).stdout
blocks, cur = [], None
for row in csv.reader(io.StringIO(out)):
    if not row:
        continue
    if row[0] == "Kernel Name":
        cur = {"name": row[1], "hdr": None, "rows": []}
        blocks.append(cur)
    elif cur is not None and cur["hdr"] is None:
        cur["hdr"] = row
    elif cur is not None:
        cur["rows"].append(row)
for b in blocks:
    H = b["hdr"]
    si = H.index("# Samples")
    src = H.index("Source")
    stall_cols = [i for i, h in enumerate(H) if h.startswith("stall_") and "Not Issued" not in h]
    ex = H.index("Instructions Executed")
    total = sum(int(r[si] or 0) for r in b["rows"])
    print(f"== {b['name'][:100]}  total samples {total}")
    ranked = sorted(enumerate(b["rows"]), key=lambda t: -int(t[1][si] or 0))[:top]
    for idx, r in ranked:
        s = int(r[si] or 0)
        reasons = sorted(((int(r[i] or 0), H[i][6:]) for i in stall_cols), reverse=True)[:3]
        rs = " ".join(f"{n}:{c}" for c, n in reasons if c)
        print(f"{idx:5d} {100.0 * s / max(total, 1):5.1f}%  exec {r[ex]:>9}  {r[src][:70]:70s} {rs}")
