#!/usr/bin/env python
"""Summarise an .ncu-rep (ncu --set full) into a small JSON + markdown table for profiles/.
usage: python tools/ncu_summary.py gpurun_out/r01_tile_full.ncu-rep profiles/r01_tile_full"""
import csv, io, json, subprocess, sys

KEYS = {
    "gpu__time_duration.sum": "duration_ms",
    "dram__bytes_read.sum": "dram_read_MB",
    "dram__bytes_write.sum": "dram_write_MB",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2_pct",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct",
    "l1tex__t_sector_hit_rate.pct": "l1_hit_pct",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_slot_util_pct",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active": "alu_pipe_pct",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active": "fma_pipe_pct",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active": "lsu_pipe_pct",
    "smsp__thread_inst_executed_per_inst_executed.ratio": "active_lanes_per_inst",
    "smsp__inst_executed.sum": "warp_instructions",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "achieved_occupancy_pct",
    "launch__registers_per_thread": "registers",
    "smsp__warps_eligible.avg.per_cycle_active": "eligible_warps_per_cycle",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum": "global_load_sectors",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld_lookup_miss.sum": "l1_miss_sectors",
    "lts__t_sectors_srcunit_tex_op_read.sum": "l2_read_sectors",
    "lts__t_sectors_srcunit_tex_op_write.sum": "l2_write_sectors",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio": "stall_wait",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio": "stall_long_scoreboard",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio": "stall_math_pipe_throttle",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio": "stall_not_selected",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio": "stall_short_scoreboard",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio": "stall_branch_resolving",
}
rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
res = []
for r in rows[2:]:
    d = {"kernel": r[hdr.index("Kernel Name")][:60], "grid": r[hdr.index("Grid Size")], "block": r[hdr.index("Block Size")]}
    for k, name in KEYS.items():
        if k in hdr:
            v = float(r[hdr.index(k)].replace(",", ""))
            u = units[hdr.index(k)]
            if name.endswith("_MB"):
                v *= {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1.0)
            if name == "duration_ms":
                v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(u, 1.0)
            d[name] = round(v, 3)
    d["dram_traffic_MB"] = round(d.get("dram_read_MB", 0) + d.get("dram_write_MB", 0), 3)
    res.append(d)
json.dump(res, open(out + ".json", "w"), indent=1)
cols = ["kernel", "duration_ms", "dram_traffic_MB", "dram_pct", "l2_pct", "l1_hit_pct", "l2_hit_pct", "issue_slot_util_pct", "alu_pipe_pct",
        "fma_pipe_pct", "active_lanes_per_inst", "achieved_occupancy_pct", "registers", "warp_instructions", "l1_miss_sectors",
        "stall_wait", "stall_long_scoreboard", "stall_math_pipe_throttle", "stall_not_selected"]
with open(out + ".md", "w") as f:
    f.write("| " + " | ".join(cols) + " |\n|" + "---|" * len(cols) + "\n")
    for d in res:
        f.write("| " + " | ".join(str(d.get(c, "")) for c in cols) + " |\n")
print(open(out + ".md").read())
