"""Summarise an ncu report (--set full) into the small JSON kept under profiles/:  python tools/ncu_summary.py rep.ncu-rep out.json "source note"
Run here (CPU box): `ncu -i` only reads the report."""
import csv
import io
import json
import subprocess
import sys

KEEP = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "gpu__time_duration.sum", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__average_warp_latency_per_inst_issued.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio"]

rep, out, note = sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else ""
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
kernels = []
for r in rows[2:]:
    k = {}
    for name in KEEP:
        if name in hdr:
            i = hdr.index(name)
            k[name] = (r[i] + (" " + units[i] if units[i] and name != "Kernel Name" else "")).strip()
    kernels.append(k)
json.dump({"source": note, "kernels": kernels}, open(out, "w"), indent=1)
for k in kernels:
    print(k["Kernel Name"][:60], k.get("gpu__time_duration.sum"), k.get("dram__bytes_read.sum"), k.get("dram__bytes_write.sum"))
