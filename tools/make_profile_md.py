"""Summaries for profiles/: python tools/make_profile_md.py <tag>
   gpurun_out/launches_<tag>.csv      -> profiles/<tag>_bench_launches.md   (kernel shares of the bench command)
   gpurun_out/prof_<tag>_*.ncu-rep    -> profiles/<tag>_<name>_ncu_full.md  (selected ncu --set full metrics per launch)
   and profiles/<tag>_traffic.json    (dram bytes of the dominant kernel, read by bench.py for roofline.traffic)"""
import collections
import csv
import glob
import json
import os
import re
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
# optional: directory holding launches_<tag>.csv / prof_<tag>_*.ncu-rep and directory to write the summaries to (on the GPU box the
# reports stay in /tmp -- they are tens of MB each -- and only the summaries travel back through gpurun_out/)
GO = sys.argv[2] if len(sys.argv) > 2 else os.path.join(REPO, "gpurun_out")
OUT = sys.argv[3] if len(sys.argv) > 3 else os.path.join(REPO, "profiles")
os.makedirs(OUT, exist_ok=True)

COLS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size"]


def short(name):
    name = re.sub(r"\(anonymous namespace\)::|<unnamed>::|void ", "", name)
    return name.split("(")[0][:70]


def launches():
    path = os.path.join(GO, f"launches_{tag}.csv")
    if not os.path.exists(path):
        return
    rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        v = float(r[vi].replace(",", ""))
        v = v / 1e3 if r[ui] in ("ns", "nsecond") else v * 1e3 if r[ui] in ("ms", "msecond") else v * 1e6 if r[ui] in ("s", "second") else v
        k = short(r[ki])
        a = agg.setdefault(k, [0.0, 0])
        a[0] += v
        a[1] += 1
    total = sum(a[0] for a in agg.values())
    n = sum(a[1] for a in agg.values())
    with open(os.path.join(OUT, f"{tag}_bench_launches.md"), "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none: python bench.py --steps 2 --warmup 3 --no-cpu-baseline (gpurun_out/launches_{tag}.csv)\n\n")
        f.write(f"{n} launches, {total / 1e3:.3f} ms total device time (cold-cache, serialised: compare SHARES, not absolutes).\n\n")
        f.write("| share | ms | launches | kernel |\n|---|---|---|---|\n")
        for k, (us, c) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
            f.write(f"| {us / total * 100:.2f}% | {us / 1e3:.3f} | {c} | {k} |\n")
    print("wrote launches", n, total / 1e3)


def full(rep):
    name = os.path.basename(rep)[len(f"prof_{tag}_"):-len(".ncu-rep")]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = [hdr.index(c) for c in COLS if c in hdr]
    ki = hdr.index("Kernel Name")
    recs = []
    with open(os.path.join(OUT, f"{tag}_{name}_ncu_full.md"), "w") as f:
        f.write(f"# ncu --set full --clock-control none (gpurun_out/{os.path.basename(rep)})\n\nPer-launch values (cold-cache, serialised replays).\n\n")
        f.write("| # | kernel | " + " | ".join(f"{hdr[i]} [{units[i]}]" for i in idx) + " |\n")
        f.write("|---|---|" + "---|" * len(idx) + "\n")
        for n, r in enumerate(rows[2:]):
            f.write(f"| {n} | {short(r[ki])} | " + " | ".join(r[i] for i in idx) + " |\n")
            recs.append({"kernel": short(r[ki]), **{hdr[i]: (r[i], units[i]) for i in idx}})
    print("wrote", name, len(recs))
    return name, recs


def to_bytes(v, u):
    x = float(v.replace(",", ""))
    return x * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]


launches()
traffic = {}
for rep in sorted(glob.glob(os.path.join(GO, f"prof_{tag}_*.ncu-rep"))):
    name, recs = full(rep)
    if name.startswith("mwt"):
        frames = int(name[3:])
        # the multiscale conv is the longest launch of the capture
        def dur(r):
            v, u = r["gpu__time_duration.sum"]
            return float(v.replace(",", "")) * {"us": 1e-3, "ms": 1.0, "ns": 1e-6, "s": 1e3}.get(u, 1.0)
        top = max(recs, key=dur)
        traffic[f"multiscale_{frames}_frames"] = {
            "dram_bytes": to_bytes(*top["dram__bytes_read.sum"]) + to_bytes(*top["dram__bytes_write.sum"]),
            "frames": frames, "source": f"profiles/{tag}_{name}_ncu_full.md (longest launch)", "ms_under_ncu": dur(top)}
    if name == "dwt256":
        r = recs[0]
        traffic["dwt3_256_frames"] = {"dram_bytes": to_bytes(*r["dram__bytes_read.sum"]) + to_bytes(*r["dram__bytes_write.sum"]),
                                      "frames": 256, "source": f"profiles/{tag}_{name}_ncu_full.md"}
if traffic:
    json.dump(traffic, open(os.path.join(OUT, f"{tag}_traffic.json"), "w"), indent=1)
    print(traffic)
