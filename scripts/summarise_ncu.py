"""Turn ncu output into the markdown summaries kept under profiles/.

    python scripts/summarise_ncu.py launches <launches.csv> <title>         (ncu --metrics gpu__time_duration.sum --csv log)
    python scripts/summarise_ncu.py full <report.ncu-rep> <title> [regex]   (ncu --set full capture; needs ncu on PATH)
"""
import collections
import csv
import io
import re
import subprocess
import sys

FULL_METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
]


def short(name):
    name = re.sub(r"\(.*", "", name)
    name = name.replace("void ", "").replace("<unnamed>::", "")
    return name.strip()


def launches(path, title):
    text = open(path).read()
    start = text.index('"ID"')
    rows = list(csv.DictReader(io.StringIO(text[start:])))
    agg = collections.OrderedDict()
    for r in rows:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        us = v / 1e3 if unit in ("nsecond", "ns") else (v * 1e3 if unit in ("msecond", "ms") else v)
        k = short(r["Kernel Name"])
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += us
    total = sum(a[1] for a in agg.values())
    print(f"# {title}\n")
    print("Per-launch times under ncu are cold-cache and serialised (the latency-bound Louvain steps and the small fp64 "
          "kernels are inflated most): compare SHARES.\n")
    print("| kernel | launches | total ms | mean us | share |\n|---|---:|---:|---:|---:|")
    for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| {k} | {n} | {us / 1e3:.3f} | {us / n:.1f} | {100 * us / total:.1f}% |")


def full(path, title, pattern=None):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    print(f"# {title}\n")
    for r in rows[2:]:
        name = r[idx["Kernel Name"]]
        if pattern and not re.search(pattern, name):
            continue
        print(f"## {short(name)}")
        for m in FULL_METRICS:
            if m in idx:
                print(f"- {m}: {r[idx[m]]} {units[idx[m]]}")
        print()


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3])
    else:
        full(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else None)
