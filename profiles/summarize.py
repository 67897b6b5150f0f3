#!/usr/bin/env python
"""Turn the ncu outputs brought back in gpurun_out/ into the small text summaries committed here.

usage: python profiles/summarize.py <tag>      (reads gpurun_out/launches_<tag>.csv and gpurun_out/prof_<tag>.ncu-rep,
                                                writes profiles/<tag>_launches.txt and profiles/<tag>_full.txt)
"""
import collections
import csv
import io
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
           "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "lts__t_sector_hit_rate.pct",
           "l1tex__t_sector_hit_rate.pct", "smsp__inst_executed.sum", "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_alu.sum",
           "sm__inst_executed_pipe_xu.sum", "sm__inst_executed_pipe_lsu.sum",
           "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
           "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
           "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio"]


def launches(tag):
    src = os.path.join(ROOT, "gpurun_out", f"launches_{tag}.csv")
    if not os.path.exists(src):
        return
    rows = [r for r in csv.reader(open(src)) if len(r) > 5]
    hdr = [i for i, r in enumerate(rows) if r[0] == "ID"][0]
    h, rows = rows[hdr], rows[hdr + 1:]
    ki, vi = h.index("Kernel Name"), h.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows:
        agg.setdefault(r[ki].split("(")[0], []).append(float(r[vi].replace(",", "")))
    tot = sum(sum(v) for v in agg.values())
    out = [f"# ncu --metrics gpu__time_duration.sum --clock-control none  ({os.path.basename(src)}, {len(rows)} launches)",
           "# per-launch times are cold-cache and serialised: compare SHARES, not absolutes",
           f"{'kernel':34s} {'launches':>8s} {'total_us':>11s} {'mean_us':>9s} {'share':>6s}"]
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        out.append(f"{k:34s} {len(v):8d} {sum(v) / 1e3:11.1f} {sum(v) / len(v) / 1e3:9.2f} {sum(v) / tot:6.3f}")
    open(os.path.join(ROOT, "profiles", f"{tag}_launches.txt"), "w").write("\n".join(out) + "\n")
    print("\n".join(out))


def full(tag):
    rep = os.path.join(ROOT, "gpurun_out", f"prof_{tag}.ncu-rep")
    if not os.path.exists(rep):
        return
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    h, units, rows = rows[0], rows[1], rows[2:]
    cols = [h.index("Kernel Name")] + [h.index(m) for m in METRICS if m in h]
    out = [f"# ncu --set full --clock-control none --import-source on  ({os.path.basename(rep)}); one block per captured launch"]
    for r in rows:
        out.append("")
        for c in cols:
            out.append(f"{h[c]:62s} {r[c][:60]:>24s} {units[c]}")
    open(os.path.join(ROOT, "profiles", f"{tag}_full.txt"), "w").write("\n".join(out) + "\n")
    # per-kernel DRAM traffic per launch (mean over the captured launches) -> profiles/traffic.json, read by bench.py
    import json
    ki, ri, wi = h.index("Kernel Name"), h.index("dram__bytes_read.sum"), h.index("dram__bytes_write.sum")
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    acc = {}
    for r in rows:
        name = r[ki].split("(")[0]
        acc.setdefault(name, []).append(float(r[ri]) * scale[units[ri]] + float(r[wi]) * scale[units[wi]])
    tj = {"source": f"profiles/{tag}_full.txt (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum, mean per launch)"}
    for name, v in acc.items():
        tj[name] = {"dram_bytes_per_launch": sum(v) / len(v), "launches": len(v)}
    # cache hit rates of the ray caster's look-ups (BASELINE north_star asks for the L2 hit rate of the probes)
    l1, l2 = h.index("l1tex__t_sector_hit_rate.pct"), h.index("lts__t_sector_hit_rate.pct")
    rc = [r for r in rows if "raycast_kernel" in r[ki]]
    if rc:
        tj["raycast_kernel_hit_rates"] = {"l1tex_sector_hit_pct": sum(float(r[l1]) for r in rc) / len(rc),
                                          "l2_sector_hit_pct": sum(float(r[l2]) for r in rc) / len(rc), "launches": len(rc)}
    json.dump(tj, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
    print("\n".join(out[:20]))


if __name__ == "__main__":
    launches(sys.argv[1])
    full(sys.argv[1])
