"""Key metrics of an `ncu --page raw --csv` export, one block per kernel launch."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
want = ["Kernel Name", "gpu__time_duration.sum", "launch__registers_per_thread", "smsp__inst_executed.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "sm__inst_executed_pipe_lsu.sum", "smsp__inst_executed_op_shared_ld.sum", "sm__warps_active.avg.pct_of_peak_sustained_active"]
for r in rows[2:]:
    print("----")
    for w in want:
        if w in idx:
            print(f"{w} = {r[idx[w]]} {units[idx[w]]}")
    for h in hdr:
        if "issue_stalled" in h and h.endswith("per_issue_active.ratio"):
            try:
                if float(r[idx[h]]) > 0.25:
                    print("    stall", h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), r[idx[h]])
            except ValueError:
                pass
