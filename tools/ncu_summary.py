"""Print the handful of ncu metrics we track from a .ncu-rep (run here, no GPU needed)."""
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__inst_executed_op_shared_ld.sum", "launch__grid_size", "launch__block_size",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "l1tex__throughput.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    print("##", r[hdr.index("Kernel Name")][:90])
    for k in hdr:
        if k in KEYS or ("issue_stalled" in k and k.endswith("per_issue_active.ratio")):
            v = r[hdr.index(k)]
            try:
                if "stalled" in k and float(v) < 0.15:
                    continue
            except ValueError:
                pass
            print(f"  {k:85s} {v} {units[hdr.index(k)]}")
