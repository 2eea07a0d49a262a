"""Summarise an .ncu-rep (ncu --set full) into a small text file that can be committed.
Usage: python profiles/summarize_ncu.py gpurun_out/x.ncu-rep profiles/x.txt ["note"]"""
import collections
import csv
import io
import re
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
note = sys.argv[3] if len(sys.argv) > 3 else ""
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
KEYS = [
    "Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "smsp__maximum_warps_per_active_cycle_pct", "sm__warps_active.avg.pct_of_peak_sustained_active", "gpu__time_duration.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "sm__cycles_elapsed.max",
]
lines = [f"# ncu summary of {rep}", f"# {note}", ""]
for k, r in enumerate(data):
    lines.append(f"## launch {k}")
    for key in KEYS:
        if key in hdr:
            i = hdr.index(key)
            lines.append(f"{key:72s} {r[i]} {units[i]}")
    lines.append("")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
blocks, cur = [], None
for r in csv.reader(io.StringIO(src)):
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}
        blocks.append(cur)
    elif cur is not None:
        cur["rows"].append(r)
for k, b in enumerate(blocks):
    if not b["rows"]:
        continue
    h, d = b["rows"][0], b["rows"][1:]
    iS, iI = h.index("Source"), h.index("Instructions Executed")
    stall = [i for i, x in enumerate(h) if x.startswith("stall_") and "Not Issued" not in x]
    tot = sum(float(r[iI] or 0) for r in d)
    ops, st = collections.Counter(), collections.Counter()
    for r in d:
        m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[iS])
        ops[(m.group(2).split(".")[0] if m else "?")] += float(r[iI] or 0)
        for i in stall:
            st[h[i]] += float(r[i] or 0)
    lines.append(f"## launch {k}: SASS mix (share of {tot:.0f} warp instructions) and stall samples")
    lines.append("  " + "  ".join(f"{o}:{c / tot * 100:.1f}%" for o, c in ops.most_common(14)))
    ssum = sum(st.values()) or 1
    lines.append("  " + "  ".join(f"{o}:{c / ssum * 100:.1f}%" for o, c in st.most_common(8)))
    lines.append("")
open(out, "w").write("\n".join(lines) + "\n")
print("\n".join(lines))
