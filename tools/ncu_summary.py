"""Summarise an .ncu-rep (raw page) into the handful of numbers DESIGN.md / bench.py quote.
    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [kernel-regex] > profiles/xyz.txt"""
import csv, io, re, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ["Kernel Name", "Block Size", "Grid Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor_op_dmma.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_op_dmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.max",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio"]
idx = {k: i for i, k in enumerate(hdr)}
extra = [k for k in hdr if re.search(r"dmma|tensor", k) and k not in want][:12]
for r in rows[2:]:
    print("=" * 100)
    for k in want + extra:
        if k in idx:
            print(f"{k:95s} {r[idx[k]]:>18s} {units[idx[k]]}")
