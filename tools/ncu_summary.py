"""Turn ncu artefacts brought back in gpurun_out/ into the small, tracked summaries under profiles/.

    python tools/ncu_summary.py full   gpurun_out/prof.ncu-rep      profiles/r01_mwt_kernels_full.md
    python tools/ncu_summary.py launch gpurun_out/launches.csv      profiles/r01_bench_launches.md
"""
import collections
import csv
import re
import subprocess
import sys

METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
           "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "launch__registers_per_thread",
           "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size"]


def short(name):
    name = re.sub(r"\(.*$", "", name).replace("(anonymous namespace)::", "").replace("<unnamed>::", "")
    return name.replace("void ", "").strip()


def full(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    cols = [(m, hdr.index(m)) for m in METRICS if m in hdr]
    ki = hdr.index("Kernel Name")
    with open(out, "w") as f:
        f.write(f"# ncu --set full --clock-control none ({rep})\n\nPer-launch values (cold-cache, serialised replays).\n\n")
        f.write("| # | kernel | " + " | ".join(f"{m} [{units[i]}]" for m, i in cols) + " |\n")
        f.write("|---|---|" + "---|" * len(cols) + "\n")
        for n, r in enumerate(rows[2:]):
            f.write(f"| {n} | {short(r[ki])} | " + " | ".join(r[i] for _, i in cols) + " |\n")
    print("wrote", out)


def launch(path, out):
    lines = [l for l in open(path) if l.startswith('"')]
    rd = csv.reader(lines)
    hdr = next(rd)
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.defaultdict(lambda: [0, 0.0])
    tot = 0.0
    for r in rd:
        v = float(r[vi].replace(",", "")) * {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}[r[ui]]
        name = short(r[ki])
        if "elementwise_kernel" in name:
            name = name.split("<")[0] + " (torch)"
        name = name[:90]
        agg[name][0] += 1
        agg[name][1] += v
        tot += v
    with open(out, "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none ({path})\n\n")
        f.write(f"{sum(a[0] for a in agg.values())} launches, {tot / 1e6:.3f} ms total device time "
                "(cold-cache, serialised: compare SHARES, not absolutes).\n\n| share | ms | launches | kernel |\n|---|---|---|---|\n")
        for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| {100 * t / tot:.2f}% | {t / 1e6:.3f} | {c} | {k} |\n")
    print("wrote", out)


if __name__ == "__main__":
    {"full": full, "launch": launch}[sys.argv[1]](sys.argv[2], sys.argv[3])
