"""Condense an `ncu --page raw --csv` export into the per-launch metrics the round summary cites."""
import csv, sys
COLS = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor_pipe_%"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_%"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_%"),
        ("lts__t_sector_hit_rate.pct", "l2_hit_%"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_%"),
        ("launch__registers_per_thread", "regs"),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall_long_sb"),
        ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall_short_sb"),
        ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall_barrier"),
        ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall_wait")]
rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
print("# from %s (ncu --set full --clock-control none; cold-cache replays, per launch)" % sys.argv[1].split("/")[-1])
for r in data:
    name = r[hdr.index("Kernel Name")].replace("(GemmArgs)", "").replace("void ", "")
    name = name.split("(")[0] if "<" not in name else name
    out = ["%-46s grid %-14s" % (name[:46], r[hdr.index("Grid Size")].replace(" ", ""))]
    for col, short in COLS:
        if col in hdr:
            i = hdr.index(col)
            v = r[i]
            try:
                v = "%.3g" % float(v)
            except ValueError:
                pass
            u = units[i] if short in ("time", "dram_rd", "dram_wr") else ""
            out.append("%s=%s%s" % (short, v, u))
    print("  ".join(out))
