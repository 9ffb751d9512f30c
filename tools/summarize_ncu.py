"""Summarise ncu outputs into the small text files committed under profiles/.

  python tools/summarize_ncu.py launches gpurun_out/x_launches.csv      # per-kernel device time shares
  python tools/summarize_ncu.py full gpurun_out/x.ncu-rep               # key metrics of a --set full capture
"""
import csv
import io
import re
import subprocess
import sys
from collections import defaultdict

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__cycles_elapsed.max",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "inst_executed", "smsp__warps_eligible.avg.per_cycle_active"]


def short(name):
    name = re.sub(r"\(.*", "", name)
    name = re.sub(r"^void ", "", name)
    return name[:70]


def launches(path):
    rows = [r for r in csv.reader(open(path, errors="replace")) if r]
    start = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[start]
    kn, mv, mn = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
    mu = hdr.index("Metric Unit")
    tot = defaultdict(lambda: [0, 0.0])
    for r in rows[start + 1:]:
        if len(r) <= mv or r[mn] != "gpu__time_duration.sum":
            continue
        t = float(r[mv].replace(",", ""))
        unit = r[mu]
        t_us = t / 1e3 if unit in ("ns", "nsecond") else t if unit in ("us", "usecond") else t * 1e3
        e = tot[short(r[kn])]
        e[0] += 1
        e[1] += t_us
    total = sum(v[1] for v in tot.values())
    print(f"# {path}: {sum(v[0] for v in tot.values())} launches, {total / 1e3:.2f} ms of device time (ncu: cold-cache, serialised)")
    print(f"{'kernel':72s} {'launches':>8s} {'total ms':>10s} {'share':>7s}")
    for k, (c, t) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:72s} {c:8d} {t / 1e3:10.3f} {100 * t / total:6.1f}%")


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print(f"## {short(r[hdr.index('Kernel Name')])}")
        for k in KEYS:
            if k in hdr:
                print(f"  {k:90s} {r[hdr.index(k)]:>18s} {units[hdr.index(k)]}")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
