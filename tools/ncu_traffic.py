"""Build profiles/<round>_conv_dram_traffic.md (stdout) and, with a second argument, profiles/roofline_traffic.json from the csv of
  LT_ITERS=1 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,\
sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,\
gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed --clock-control none --csv --log-file X.csv \
  -k regex:"conv_tc_kernel|fir|act_bwd|torgb" python tools/layer_times.py 1024 20 tf32
usage: python tools/ncu_traffic.py X.csv [profiles/roofline_traffic.json] > profiles/r02_conv_dram_traffic.md"""
import csv, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10 and r[0].isdigit()]
per = collections.OrderedDict()
for r in rows:
    per.setdefault(int(r[0]), {"name": r[4]})[r[-3]] = (float(r[-1].replace(",", "")), r[-2])
def val(d, k, scale_to=None):
    v, u = d.get(k, (0.0, ""))
    if scale_to == "us":
        return v / 1e3 if u in ("ns", "nsecond") else (v if u in ("us", "usecond") else v * 1e3 if u in ("ms", "msecond") else v)
    if scale_to == "MB":
        return {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1e-6) * v
    return v
out, conv_bytes, conv_n = [], 0.0, 0
for i, (k, d) in enumerate(per.items()):
    name = d["name"].split("(")[0].replace("void ", "").replace("lfp::", "")
    t = val(d, "gpu__time_duration.sum", "us")
    rd, wr = val(d, "dram__bytes_read.sum", "MB"), val(d, "dram__bytes_write.sum", "MB")
    out.append(f"| {i} | `{name}` | {t:.1f} | {rd:.1f} | {wr:.1f} | {val(d, 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):.1f} | "
               f"{val(d, 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'):.1f} | {val(d, 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed'):.1f} |")
    if "conv_tc_kernel" in name:
        conv_bytes += (rd + wr) * 1e6; conv_n += 1
print("# ncu metrics of every conv / FIR / act_bwd / ToRGB launch of one forward+backward, 1024 px, B=20, tf32 path (round 2)\n")
if len(sys.argv) > 2:
    import json
    json.dump({"attribution_1024px_n20_mse/tf32": {"dram_bytes_per_conv_launch": conv_bytes / max(conv_n, 1), "conv_launches": conv_n,
                                                    "source": "profiles/r02_conv_dram_traffic.md (ncu dram__bytes_read.sum + dram__bytes_write.sum per conv_tc_kernel launch, one forward+backward at B = 20)"}},
              open(sys.argv[2], "w"), indent=1)
print("Command: see tools/ncu_traffic.py (serialised, cold-cache per launch: compare shares, not absolutes).\n")
print(f"Average DRAM bytes (read + write) per conv launch: **{conv_bytes / max(conv_n, 1):.4g}** over {conv_n} conv launches (`roofline.traffic` in bench.py).\n")
print("| # | kernel | time us | dram rd MB | dram wr MB | dram % | tensor pipe % | L1/smem % |\n|---|---|---|---|---|---|---|---|")
print("\n".join(out))
