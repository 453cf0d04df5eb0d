"""Tabulate key metrics of every kernel in an `ncu --page raw --csv` export."""
import csv
import sys

KEYS = ["gpu__time_duration.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "smsp__inst_executed.sum", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct",
        "lts__t_sector_hit_rate.pct",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__cycles_elapsed.max",
        "sm__cycles_elapsed.avg.per_second"]


def main(path):
    rows = list(csv.reader(open(path)))
    hdr = rows[0]
    names = [r[hdr.index("Kernel Name")].replace("void ", "").replace("(LoopArgs)", "").replace("loopKernel", "")[:18]
             for r in rows[2:]]
    print(" " * 62, *[n.rjust(18) for n in names])
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            print(k.replace("smsp__average_warps_issue_stalled_", "stall_").replace("_per_issue_active.ratio", "")[:62]
                  .ljust(62), *[r[i][:12].rjust(18) for r in rows[2:]], rows[1][i][:8])


if __name__ == "__main__":
    main(sys.argv[1])
