#!/usr/bin/env python
"""Per-kernel summary of an `ncu --set full` report: duration, instructions, issue utilisation, DRAM traffic per launch.
usage: ncu_traffic.py report.ncu-rep seqs_per_launch out.json [out.md]
The JSON is what bench.py reads for `roofline.traffic` (dram__bytes_read.sum + dram__bytes_write.sum per launch)."""
import csv, json, subprocess, sys

rep, seqs, out_json = sys.argv[1], int(sys.argv[2]), sys.argv[3]
out_md = sys.argv[4] if len(sys.argv) > 4 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL).stdout.decode()
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}
U = {h: units[i] for i, h in enumerate(hdr)}

def num(r, name, default=0.0):
    try:
        return float(r[col[name]].replace(",", ""))
    except (KeyError, ValueError):
        return default

def to_bytes(v, unit):
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)

def to_us(v, unit):
    return v * {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(unit, 1)

kern = {}
for r in rows[2:]:
    if len(r) != len(hdr):
        continue
    name = r[col["Kernel Name"]].split("(")[0].replace("void ", "")
    if name in kern:
        continue                          # first captured launch of each kernel
    kern[name] = {
        "duration_us": to_us(num(r, "gpu__time_duration.sum"), U["gpu__time_duration.sum"]),
        "dram_read_bytes": to_bytes(num(r, "dram__bytes_read.sum"), U["dram__bytes_read.sum"]),
        "dram_write_bytes": to_bytes(num(r, "dram__bytes_write.sum"), U["dram__bytes_write.sum"]),
        "warp_instructions": num(r, "smsp__inst_executed.sum"),
        "issue_active_pct": num(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
        "warps_active_pct": num(r, "sm__warps_active.avg.pct_of_peak_sustained_active"),
        "registers": num(r, "launch__registers_per_thread"),
        "l1_hit_pct": num(r, "l1tex__t_sector_hit_rate.pct"),
        "l2_hit_pct": num(r, "lts__t_sector_hit_rate.pct"),
        "lsu_wavefront_pct": num(r, "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"),
        "grid": r[col["Grid Size"]] if "Grid Size" in col else "",
    }
    kern[name]["traffic_bytes"] = kern[name]["dram_read_bytes"] + kern[name]["dram_write_bytes"]
json.dump({"source": rep, "seqs_per_launch": seqs, "kernels": kern}, open(out_json, "w"), indent=1)
lines = ["| kernel | duration us | warp instr | issue active % | warps active % | regs | LSU wavefronts % | L1 hit % | L2 hit % | dram read MB | dram write MB |", "|---|---|---|---|---|---|---|---|---|---|---|"]
for k, v in sorted(kern.items(), key=lambda kv: -kv[1]["duration_us"]):
    lines.append("| %s | %.1f | %.1fM | %.1f | %.1f | %d | %.1f | %.1f | %.1f | %.1f | %.1f |" % (
        k, v["duration_us"], v["warp_instructions"] / 1e6, v["issue_active_pct"], v["warps_active_pct"], v["registers"], v["lsu_wavefront_pct"],
        v["l1_hit_pct"], v["l2_hit_pct"], v["dram_read_bytes"] / 1e6, v["dram_write_bytes"] / 1e6))
print("\n".join(lines))
if out_md:
    open(out_md, "w").write("\n".join(lines) + "\n")
