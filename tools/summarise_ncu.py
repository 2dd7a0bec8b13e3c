"""Turn ncu CSV exports into the small summaries kept under profiles/.

  python tools/summarise_ncu.py raw  <ncu --page raw --csv file>   <out.csv>  ["comment"]
  python tools/summarise_ncu.py list <ncu --metrics gpu__time_duration.sum --csv log> <out_summary.csv> ["comment"]
"""
import csv
import sys
from collections import OrderedDict

COLS = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "gpu__time_duration.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "sm__cycles_active.avg"]


def raw(src, dst, comment):
    rows = [r for r in csv.reader(open(src)) if r]
    start = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr, units, body = rows[start], rows[start + 1], rows[start + 2:]
    idx = [hdr.index(c) for c in COLS if c in hdr]
    with open(dst, "w", newline="") as f:
        if comment:
            f.write('"# %s"\n' % comment.replace('"', "'"))
        w = csv.writer(f)
        w.writerow([hdr[i] for i in idx])
        w.writerow([units[i] for i in idx])
        for r in body:
            w.writerow([r[i] for i in idx])


def launch_list(src, dst, comment):
    rows = [r for r in csv.reader(open(src)) if r]
    start = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr, body = rows[start], rows[start + 1:]
    kn, mv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = OrderedDict()
    for r in body:
        if len(r) <= mv:
            continue
        name = r[kn].split("(")[0].replace("void ", "")[:90]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += float(r[mv].replace(",", ""))
    tot = sum(a[1] for a in agg.values())
    with open(dst, "w", newline="") as f:
        if comment:
            f.write('"# %s"\n' % comment.replace('"', "'"))
        w = csv.writer(f)
        w.writerow(["kernel", "launches", "total_ns", "share"])
        for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            w.writerow([name, n, t, "%.4f" % (t / tot)])


if __name__ == "__main__":
    mode, src, dst = sys.argv[1:4]
    comment = sys.argv[4] if len(sys.argv) > 4 else ""
    (raw if mode == "raw" else launch_list)(src, dst, comment)
