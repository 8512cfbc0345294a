"""ncu `--page raw --csv` export -> markdown table of the metrics the roofline discussion uses, one column per launch.
   python profiles/scripts/summarise_raw.py <raw.csv> [kernel-name-filter]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
flt = sys.argv[2] if len(sys.argv) > 2 else ""
ki = hdr.index("Kernel Name")
data = [r for r in data if flt in r[ki]]
M = [("gpu__time_duration.sum", "duration"), ("launch__registers_per_thread", "registers / thread"), ("launch__grid_size", "grid"),
     ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
     ("sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "heavy FMA pipe active %"),
     ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe %"),
     ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe %"),
     ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
     ("smsp__inst_executed.sum", "warp instructions"),
     ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
     ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
     ("l1tex__t_sector_hit_rate.pct", "L1 hit %"), ("lts__t_sector_hit_rate.pct", "L2 hit %"),
     ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "L1TEX wavefronts %"),
     ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %")]
for name in hdr:
    if name.startswith("smsp__average_warps_issue_stalled_") and name.endswith("_per_issue_active.ratio"):
        M.append((name, "stall: " + name[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
print("| metric | " + " | ".join("#%d" % (i + 1) for i in range(len(data))) + " |")
print("|---|" + "---|" * len(data))
print("| kernel | " + " | ".join(r[ki].split("(")[0][-40:] for r in data) + " |")
for key, label in M:
    if key not in hdr:
        continue
    i = hdr.index(key)
    vals = []
    for r in data:
        try:
            v = float(r[i].replace(",", ""))
        except ValueError:
            v = None
        vals.append(v)
    if label.startswith("stall") and all((v or 0) < 0.1 for v in vals):
        continue
    def fmt(v):
        if v is None:
            return "-"
        if abs(v) >= 1e6:
            return "%.3g" % v
        return "%.3f" % v if abs(v) < 100 else "%.1f" % v
    print("| %s [%s] | " % (label, units[i]) + " | ".join(fmt(v) for v in vals) + " |")
