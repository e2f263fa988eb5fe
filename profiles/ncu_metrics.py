"""Print the key metrics of an `ncu --page raw --csv` export (one row per profiled launch)."""
import csv
import sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__grid_size',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.sum', 'sm__inst_executed_pipe_alu.sum', 'sm__inst_executed_pipe_lsu.sum',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__thread_inst_executed_per_inst_executed.ratio',
        'l1tex__t_bytes.sum', 'lts__t_bytes.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'smsp__average_warp_latency_issue_stalled']


def main(path, row=0):
    rows = list(csv.reader(open(path)))
    hdr, units, vals = rows[0], rows[1], rows[2 + row]
    for i, h in enumerate(hdr):
        if h in WANT or ('warp_issue_stalled' in h and h.endswith('per_warp_active.pct')):
            print(f"{h:90s} {units[i]:14s} {vals[i]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0)
