#!/usr/bin/env python
"""Condenses ncu output into the small CSVs kept under profiles/.

  ncu_summary.py launches <launches.csv> <out.csv>    per-kernel totals and share of the step (gpu__time_duration)
  ncu_summary.py full <report.ncu-rep> <out.csv>      the metrics DESIGN.md cites, one row per captured launch
  ncu_summary.py dram <dram.csv> <out.csv> <n_paths> <key> [<note>]
        per-kernel dram__bytes_read/write summed over EVERY launch of one render (ncu --metrics
        dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum -k regex:wf_ ...), divided by the render's
        path count; also records the total under <key> in profiles/dram_per_path.json, which bench.py READS for
        roofline.traffic (nothing is typed into bench.py).
"""
import collections
import csv
import subprocess
import sys

KEEP = [
    "Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__registers_per_thread",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__inst_executed.avg.per_cycle_active", "smsp__inst_executed.sum",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "l1tex__throughput.avg.pct_of_peak_sustained_active",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct", "dram__bytes_read.sum",
    "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__warps_eligible.avg.per_cycle_active",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
]


def launches(src, dst):
    tot, cnt = collections.Counter(), collections.Counter()
    for r in csv.reader(open(src)):
        if len(r) > 5 and "gpu__time_duration" in r[-3]:
            name = r[4].split("(")[0]
            tot[name] += float(r[-1])
            cnt[name] += 1
    s = sum(tot.values())
    with open(dst, "w") as f:
        w = csv.writer(f)
        w.writerow(["kernel", "launches", "total_us", "share_pct"])
        for k, v in tot.most_common():
            w.writerow([k, cnt[k], round(v / 1e3, 1), round(100 * v / s, 2)])
            print(f"{k:60s} n={cnt[k]:5d} {v / 1e3:10.1f} us {100 * v / s:5.1f}%")


def full(rep, dst):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    idx = {h: i for i, h in enumerate(rows[0])}
    keep = [k for k in KEEP if k in idx]
    with open(dst, "w") as f:
        w = csv.writer(f)
        w.writerow(keep)
        w.writerow([rows[1][idx[k]] for k in keep])
        for r in rows[2:]:
            w.writerow([r[idx[k]] for k in keep])
            print(" | ".join(f"{r[idx[k]]}" for k in keep[:11]))


def dram(src, dst, n_paths, key, note=""):
    import json
    import os
    n_paths = int(n_paths)
    rd, wr, tm, cnt = collections.Counter(), collections.Counter(), collections.Counter(), collections.Counter()
    for r in csv.reader(open(src)):
        if len(r) < 15 or not r[0].isdigit():
            continue
        name, metric, val = r[4].split("(")[0], r[12], float(r[14])
        if metric == "dram__bytes_read.sum":
            rd[name] += val
            cnt[name] += 1
        elif metric == "dram__bytes_write.sum":
            wr[name] += val
        elif metric == "gpu__time_duration.sum":
            tm[name] += val
    total = 0.0
    with open(dst, "w") as f:
        w = csv.writer(f)
        w.writerow([f"# {note}"])
        w.writerow(["kernel", "launches", "dram_read_bytes", "dram_write_bytes", "time_us", "bytes_per_path"])
        for k in rd:
            bpp = (rd[k] + wr[k]) / n_paths
            total += bpp
            w.writerow([k, cnt[k], int(rd[k]), int(wr[k]), round(tm[k] / 1e3, 1), round(bpp, 1)])
            print(f"{k:50s} n={cnt[k]:4d} {bpp:8.1f} B/path {tm[k] / 1e3:9.1f} us")
        w.writerow(["total", sum(cnt.values()), int(sum(rd.values())), int(sum(wr.values())),
                    round(sum(tm.values()) / 1e3, 1), round(total, 1)])
    print(f"total {total:.1f} B/path")
    reg = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "dram_per_path.json")
    table = json.load(open(reg)) if os.path.exists(reg) else {}
    table[key] = {"bytes_per_path": round(total, 1), "source": os.path.relpath(dst, os.path.dirname(os.path.dirname(reg))),
                  "per_kernel": {k: round((rd[k] + wr[k]) / n_paths, 1) for k in rd}, "note": note}
    json.dump(table, open(reg, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    {"launches": launches, "full": full, "dram": dram}[sys.argv[1]](*sys.argv[2:])
