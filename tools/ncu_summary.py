"""Key metrics of every kernel launch in an .ncu-rep (ncu --set full) as CSV rows: one file for profiles/.
Usage: ncu_summary.py report.ncu-rep > profiles/<name>_summary.csv"""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__grid_size",
        "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg", "sm__cycles_elapsed.max"]
txt = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr, units = rows[0], rows[1]
out = csv.writer(sys.stdout)
out.writerow(["kernel", "metric", "unit", "value"])
ki = hdr.index("Kernel Name")
for r in rows[2:]:
    name = r[ki].replace("lcbi::<unnamed>::", "").replace("void ", "")[:60]
    for h, u, v in zip(hdr, units, r):
        base = h.split(".TriageCompute.")[-1]
        if base in KEYS or (base.startswith("smsp__average_warps_issue_stalled") and base.endswith("ratio") and v and float(v) >= 0.3):
            out.writerow([name, base, u, v])
