#!/usr/bin/env python
"""Turn ncu output into the small CSV summaries kept under profiles/.

  ncu_summarize.py full  <report.ncu-rep> <out.csv>      selected metrics of every captured launch
  ncu_summarize.py shares <launches.csv> <out_shares.csv> [<out_launches.csv>]
                                                          per kernel/grid totals and shares of a
                                                          `--metrics gpu__time_duration.sum --csv` launch list
"""
import collections
import csv
import io
import re
import subprocess
import sys

FULL_METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "l1tex__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "smsp__inst_executed.sum",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
]


def full(rep, out):
    if rep.endswith(".csv"):  # already the raw page (exported on the GPU box)
        raw = open(rep).read()
    else:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    cols = ["Kernel Name", "Grid Size", "Block Size"] + [m for m in FULL_METRICS if m in hdr]
    idx = [hdr.index(c) for c in cols]
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(cols)
        w.writerow([units[i] for i in idx])
        for d in data:
            w.writerow([d[i] for i in idx])
    print(f"{out}: {len(data)} launches, {len(cols)} columns")


def shares(src, out_shares, out_launches=None):
    lines = [l for l in open(src) if l.startswith('"')]
    rows = list(csv.DictReader(io.StringIO("".join(lines))))
    agg = collections.OrderedDict()
    per = []
    for r in rows:
        if r["Metric Name"] != "gpu__time_duration.sum":
            continue
        name = re.sub(r"^void ", "", r["Kernel Name"])
        name = re.sub(r"\(.*\)$", "", name)[:80]
        us = float(r["Metric Value"].replace(",", "")) / (1e3 if r["Metric Unit"] == "ns" else 1.0)
        key = (name, r["Grid Size"])
        a = agg.setdefault(key, [0, 0.0])
        a[0] += 1
        a[1] += us
        per.append((r["ID"], name, r["Grid Size"], r["Block Size"], round(us, 2)))
    total = sum(a[1] for a in agg.values())
    with open(out_shares, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["kernel", "grid", "launches", "total_us", "share_pct"])
        for (name, grid), (n, us) in agg.items():
            w.writerow([name, grid, n, round(us, 1), round(100.0 * us / total, 2)])
    if out_launches:
        with open(out_launches, "w", newline="") as f:
            w = csv.writer(f)
            w.writerow(["id", "kernel", "grid", "block", "duration_us"])
            w.writerows(per)
    print(f"{out_shares}: {len(agg)} kernel/grid classes, {len(per)} launches, {total / 1e3:.2f} ms")


if __name__ == "__main__":
    if len(sys.argv) >= 4 and sys.argv[1] == "full":
        full(sys.argv[2], sys.argv[3])
    elif len(sys.argv) >= 4 and sys.argv[1] == "shares":
        shares(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else None)
    else:
        raise SystemExit(__doc__)
