"""One-screen summary of an `ncu --set full` report (the metrics DESIGN.md and bench.py cite), read on the CPU box:
    python tools/ncu_summary.py gpurun_out/x.ncu-rep "title" > profiles/x.txt"""
import csv
import subprocess
import sys

rep, title = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
WANT = ["Kernel Name", "Grid Size", "Block Size", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "gpu__time_duration.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "launch__registers_per_thread", "sm__cycles_elapsed.avg.per_second",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "lts__t_sector_hit_rate.pct", "smsp__inst_executed.sum"]
print(title)
for r in rows[2:]:
    for k in WANT:
        if k in hdr:
            i = hdr.index(k)
            print(f"{k} = {r[i]} {units[i]}")
    print("---")
