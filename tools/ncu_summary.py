"""Summarise an ncu report (read here, no GPU): python tools/ncu_summary.py report.ncu-rep > profiles/x.md"""
import csv, io, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h = rows[0]; idx = {n: i for i, n in enumerate(h)}
want = [("gpu__time_duration.sum", "time"), ("launch__grid_size", "grid"), ("launch__registers_per_thread", "regs"),
        ("dram__bytes_read.sum", "dram rd"), ("dram__bytes_write.sum", "dram wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %"),
        ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1/smem %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
        ("smsp__inst_executed.sum", "warp insts")]
print(f"ncu --set full summary of `{rep.split('/')[-1]}` (one row per profiled launch; units as reported by ncu)\n")
print("| # | kernel | " + " | ".join(n for _, n in want) + " |")
print("|---|---|" + "---|" * len(want))
for i, r in enumerate(rows[2:]):
    name = r[idx["Kernel Name"]].split("(")[0].replace("void ", "")
    cells = []
    for k, _ in want:
        v = r[idx[k]] if k in idx else ""
        u = rows[1][idx[k]] if k in idx else ""
        try:
            v = f"{float(v.replace(',', '')):.4g}"
        except ValueError:
            pass
        cells.append(f"{v} {u}".strip())
    print(f"| {i} | `{name}` | " + " | ".join(cells) + " |")
