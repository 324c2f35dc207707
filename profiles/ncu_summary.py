"""Condense an .ncu-rep into the text summary kept under profiles/.
usage: python profiles/ncu_summary.py gpurun_out/x.ncu-rep [kernel-substring] > profiles/rN_x_summary.txt"""
import csv, subprocess, sys
KEYS = ("gpu__time_duration.sum", "launch__registers_per_thread", "launch__occupancy_limit", "launch__grid_size", "launch__block_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor_subpipe_dmma", "sm__pipe_tensor_subpipe_dmma",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__t_sector_pipe_lsu_mem_local_op_ld_hit_rate.pct",
        "smsp__average_warps_issue_stalled", "smsp__sass_inst_executed_op_global", "smsp__sass_inst_executed_op_local", "smsp__sass_inst_executed_op_shared",
        "smsp__sass_thread_inst_executed_op_d", "sm__sass_thread_inst_executed_op_d", "smsp__sass_inst_executed_op_tma", "sm__throughput.avg.pct", "gpu__compute_memory_throughput")
rep = sys.argv[1]
pat = sys.argv[2] if len(sys.argv) > 2 else ""
rows = list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout.splitlines()))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")]
    if pat and pat not in name:
        continue
    print("Kernel Name =", name)
    for h, u, v in zip(hdr, units, r):
        if any(h.startswith(k) for k in KEYS) and not (h.startswith("smsp__average_warps_issue_stalled") and not h.endswith("per_issue_active.ratio")):
            print(f"{h} [{u}] = {v}")
    print()
