"""Hottest SASS lines (warp-stall samples) of the kernels in an .ncu-rep: python tools/ncu_hot.py rep [top] [kernel-index]."""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
kernels, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}
        kernels.append(cur)
    elif r and r[0] == "Address":
        cur["hdr"] = r
    elif cur is not None and r:
        cur["rows"].append(r)
sel = int(sys.argv[3]) if len(sys.argv) > 3 else None
for ki, k in enumerate(kernels):
    if sel is not None and ki != sel:
        continue
    h = k["hdr"]
    si, ie = h.index("Warp Stall Sampling (All Samples)"), h.index("Instructions Executed")
    total = sum(float(r[si] or 0) for r in k["rows"])
    print(f"== [{ki}] {k['name'][:100]}  samples={total:.0f}  sass_lines={len(k['rows'])}")
    order = sorted(range(len(k["rows"])), key=lambda i: -float(k["rows"][i][si] or 0))[:top]
    for i in sorted(order):
        r = k["rows"][i]
        print(f"  {i:5d} {float(r[si] or 0) / max(total, 1) * 100:5.1f}%  x{r[ie]:>9}  {r[1].strip()[:110]}")
