"""SASS opcode census of libewvit.so -> profiles/<tag>_sass_opcodes.md (runs without a GPU: cuobjdump only).
    python tools/sass_census.py r02"""
import collections
import os
import re
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
so = os.path.join(REPO, "efficient-wavelet-vit_b200", "ewvit", "libewvit.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
pat = re.compile(r"\b(UTC\w+(?:\.\w+)*|LDTM(?:\.\w+)*|UTMA\w+(?:\.\w+)*|UBLKCP(?:\.\w+)*|UCGABAR\w*|HMMA(?:\.\w+)*|LDSM(?:\.\w+)*|MUFU\.TANH|FFMA2|SYNCS(?:\.\w+)*|LDGSTS(?:\.\w+)*)")
total = collections.Counter()
per = collections.OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(anonymous namespace\)::|void ", "", name).split("(")[0]
        cur = per.setdefault(name, collections.Counter())
        continue
    if cur is None or "/*" not in line:
        continue
    body = line.split("*/", 1)[-1]
    for op in pat.findall(body):
        op = re.sub(r"\.(?:U32|64|128|E|SYS|STRONG|GPU|CONSTANT|BYPASS|LTC128B|ZFILL)\b", "", op) if op.startswith("LDGSTS") else op
        total[op] += 1
        cur[op] += 1
out = os.path.join(REPO, "profiles", f"{tag}_sass_opcodes.md")
with open(out, "w") as f:
    f.write(f"# SASS opcode census of `efficient-wavelet-vit_b200/ewvit/libewvit.so` ({tag})\n\n")
    f.write("`python tools/sass_census.py` = `cuobjdump -sass libewvit.so`, per kernel: tcgen05 MMA (`UTCHMMA`, `.2CTA` = cta_group::2 pairs), tcgen05 "
            "commit/barrier (`UTCBAR`), TMEM loads (`LDTM`), tiled TMA loads / stores (`UTMALDG` / `UTMASTG`), 1-D bulk copies (`UBLKCP`), "
            "cluster barrier (`UCGABAR`), legacy warp MMA (`HMMA`), `MUFU.TANH` (one-op SiLU), packed fp32x2 FMAs (`FFMA2`), asynchronous "
            "16-byte copies (`LDGSTS`).  Build of this commit, sm_100a only.\n\n## Totals\n\n| opcode | count |\n|---|---|\n")
    for op, c in total.most_common():
        f.write(f"| `{op}` | {c} |\n")
    f.write("\n## Per kernel\n\n| kernel | opcodes |\n|---|---|\n")
    for name, cnt in per.items():
        if cnt:
            f.write(f"| `{name[:110]}` | " + ", ".join(f"{op} x{c}" for op, c in cnt.most_common()) + " |\n")
print("wrote", out, sum(total.values()), "opcodes in", len(per), "kernels")
