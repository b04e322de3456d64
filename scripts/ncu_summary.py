"""Per-kernel summary of an ncu report (`ncu -i X.ncu-rep --page raw --csv | python scripts/ncu_summary.py`)."""
import csv
import sys

WANT = [
    ("gpu__time_duration.sum", "us"), ("sm__cycles_elapsed.max", "cyc"), ("sm__cycles_elapsed.avg.per_second", "GHz"),
    ("dram__bytes_read.sum", "B"), ("dram__bytes_write.sum", "B"), ("lts__t_bytes.sum", "B"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "%"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "%"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "%"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "%"),
    ("smsp__inst_executed.sum", "inst"), ("launch__registers_per_thread", ""), ("launch__grid_size", ""), ("launch__block_size", ""),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", ""), ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", ""),
    ("l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum", "B"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", ""),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", ""),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", ""),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", ""),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", ""),
    ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", ""),
    ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", ""),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", ""),
]


def main():
    rows = list(csv.reader(l for l in sys.stdin if l.startswith('"')))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        print("# kernel: %s" % r[idx["Kernel Name"]].split("(")[0])
        for name, _ in WANT:
            if name in idx:
                print("  %-78s %s %s" % (name, r[idx[name]], units[idx[name]]))


if __name__ == "__main__":
    main()
