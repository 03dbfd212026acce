"""Condense an `ncu --page raw --csv` dump (stdin or file) to the counters DESIGN.md / bench.py cite."""
import csv
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
    "launch__grid_size", "launch__block_size",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
    "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
]


def main():
    src = open(sys.argv[1]) if len(sys.argv) > 1 else sys.stdin
    rows = list(csv.reader(src))
    H, units, data = rows[0], rows[1], rows[2:]
    name_col = H.index("Kernel Name") if "Kernel Name" in H else None
    for li, r in enumerate(data):
        print(f"--- launch {li}: {r[name_col][:100] if name_col is not None else ''}")
        for k in KEYS:
            if k in H:
                i = H.index(k)
                print(f"{k:88s} {r[i]:>16s} {units[i]}")


if __name__ == "__main__":
    main()
