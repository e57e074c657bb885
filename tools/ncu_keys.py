"""Print a fixed set of ncu metrics for every launch in a report: python tools/ncu_keys.py report.ncu-rep"""
import csv, io, subprocess, sys
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw))); h = rows[0]; idx = {n: i for i, n in enumerate(h)}
keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.avg.per_cycle_elapsed", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps", "launch__occupancy_limit_blocks",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
keys += [k for k in h if "issue_stalled" in k and k.endswith("per_issue_active.ratio") and "not_issued" not in k]
for k in keys + sys.argv[2:]:
    if k in idx:
        print(f"{k[-70:]:70s}", " ".join(f"{rows[i][idx[k]]:>12s}" for i in range(2, len(rows))), rows[1][idx[k]])
