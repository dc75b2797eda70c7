"""Writes profiles/syrk_traffic.json (read by bench.py for roofline.traffic) and a section summary from an `ncu --set full`
capture of tools/syrk_only.py:  ncu -i <rep> --page raw --csv > raw.csv ; python tools/ncu_traffic.py raw.csv <label>"""
import csv, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rows = list(csv.reader(open(sys.argv[1])))
label = sys.argv[2] if len(sys.argv) > 2 else "ncu --set full"
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
r = rows[-1]                                   # last captured launch (warm)
get = lambda k: float(r[idx[k]].replace(",", ""))
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
rd = get("dram__bytes_read.sum") * scale[units[idx["dram__bytes_read.sum"]]]
wr = get("dram__bytes_write.sum") * scale[units[idx["dram__bytes_write.sum"]]]
out = {"n": 4096, "k": 4096, "dram_bytes_per_launch": rd + wr, "dram_read": rd, "dram_write": wr, "source": label}
json.dump(out, open(os.path.join(ROOT, "profiles", "syrk_traffic.json"), "w"))
want = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
        "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
for k in want:
    if k in idx:
        print(f"{k:75s} {r[idx[k]]:>20s} {units[idx[k]]}")
print(json.dumps(out))
