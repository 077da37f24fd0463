"""Digest of an .ncu-rep: key metrics + stall samples by source line (needs -lineinfo, --import-source on).
usage: python tools/ncu_digest.py report.ncu-rep [min_samples]"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
minn = int(sys.argv[2]) if len(sys.argv) > 2 else 60
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, vals = rows[0], rows[2]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum ", "dram__bytes_write.sum ", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum ",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum ", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum ", "smsp__inst_executed.sum ",
        "launch__registers_per_thread ", "smsp__average_warps_issue_stalled", "sm__cycles_active.avg ", "sm__warps_active.avg.per_cycle_active",
        "gpc__cycles_elapsed.avg.per_second", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum ", "lsu_wavefronts_mem_lgds.avg"]
for h, v in zip(hdr, vals):
    hh = h + " "
    if any(k in hh for k in want) and "Not Issued" not in h:
        try:
            if float(v) == 0.0:
                continue
        except ValueError:
            pass
        print(f"{h:95s} {v}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr, data = rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
tot = sum(int(r[ix["# Samples"]] or 0) for r in data)
print("total samples", tot)
byexec = collections.Counter()
for r in data:
    byexec[r[ix["Instructions Executed"]]] += int(r[ix["# Samples"]] or 0)
print("samples by execution count:", byexec.most_common(8))
keys = ["stall_wait", "stall_math", "stall_short_sb", "stall_long_sb", "stall_not_selected", "stall_branch_resolving", "stall_barrier", "stall_selected",
        "stall_no_inst", "stall_dispatch", "stall_mio", "stall_lg", "stall_membar"]
agg = collections.Counter()
for r in data:
    for k in keys:
        agg[k] += int(r[ix[k]] or 0)
print("stalls:", dict(agg))
for k, r in enumerate(data):
    n = int(r[ix["# Samples"]] or 0)
    if n >= minn:
        print(k, r[ix["Source"]][:64].ljust(64), str(n).rjust(5), r[ix["Instructions Executed"]].rjust(7),
              " ".join(f"{kk[6:9]}{r[ix[kk]]}" for kk in keys if int(r[ix[kk]] or 0) > 0))
