"""Print the handful of metrics we look at first from an .ncu-rep (one line per captured launch)."""
import csv, subprocess, sys
rep = sys.argv[1]
rows = list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout.splitlines()))
h = rows[0]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "launch__registers_per_thread",
        "launch__grid_size", "smsp__issue_active.avg.pct_of_peak_sustained_active", "lts__t_sectors_srcunit_tex_op_read.sum"]
for r in rows[2:]:
    print("----")
    for k in want:
        if k in h:
            print(f"{k:70s} {rows[1][h.index(k)]:12s} {r[h.index(k)]}")
    for i, n in enumerate(h):
        if "issue_stalled" in n and "per_issue_active" in n and "not_issued" not in n and float(r[i] or 0) > 0.3:
            print(f"{n:70s} {r[i]}")
