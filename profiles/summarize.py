"""Turns the scratch ncu outputs in gpurun_out/ into the small tracked summaries under profiles/.
usage: python profiles/summarize.py <tag>"""
import csv
import os
import subprocess
import sys

tag = sys.argv[1]
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = os.path.join(root, "profiles")
METRICS = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
           "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
           "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.sum", "sm__inst_executed_pipe_fma.sum",
           "sm__inst_executed_pipe_fp64.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.avg",
           "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_wait_per_warp_active.pct",
           "smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct", "smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct",
           "smsp__warp_issue_stalled_branch_resolving_per_warp_active.pct", "smsp__warp_issue_stalled_not_selected_per_warp_active.pct"]

lines = []
for rep in sorted(f for f in os.listdir(os.path.join(root, "gpurun_out")) if f.endswith(f"_{tag}.ncu-rep")):
    raw = subprocess.run(["ncu", "-i", os.path.join(root, "gpurun_out", rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    rows = [r for r in rows if len(r) > 10]
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    lines.append(f"## {rep} (ncu --set full --clock-control none)")
    for r in rows[2:]:
        lines.append("kernel: " + r[idx["Kernel Name"]].split("(")[0])
        for m in METRICS:
            if m in idx:
                lines.append(f"  {m} = {r[idx[m]]} {units[idx[m]]}")
    lines.append("")

import glob
for lp in sorted(glob.glob(os.path.join(root, "gpurun_out", f"launches_{tag}.csv")) + glob.glob(os.path.join(root, "gpurun_out", f"launches_*_{tag}.csv"))):
    agg = {}
    with open(lp, errors="ignore") as f:
        rd = csv.reader(l for l in f if l.startswith('"'))
        hdr = next(rd)
        ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
        for r in rd:
            if len(r) <= iv:
                continue
            name = r[ik].split("(")[0].replace("<unnamed>::", "")
            if "cub" in name:
                name = "cub::" + name.split("cub::")[-1].split("<")[0]
            try:
                v = float(r[iv].replace(",", ""))
            except ValueError:
                continue
            n, t = agg.get(name, (0, 0.0))
            agg[name] = (n + 1, t + v)
    tot = sum(t for _, t in agg.values())
    lines.append(f"## {os.path.basename(lp)}: per-kernel device time (gpu__time_duration.sum, ns; cold-cache, serialised => compare shares)")
    for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append(f"  {name:60s} launches={n:5d} mean_ns={t / n:10.0f} share={100 * t / tot:5.1f}%")
    lines.append("")
    # tracked copy of the launch list itself
    import shutil
    shutil.copy(lp, os.path.join(out, os.path.basename(lp)))
with open(os.path.join(out, f"summary_{tag}.txt"), "w") as f:
    f.write("\n".join(lines) + "\n")
print("\n".join(lines[-25:]))
