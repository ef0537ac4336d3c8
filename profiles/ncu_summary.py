#!/usr/bin/env python
"""Summarise `ncu -i X.ncu-rep --page raw --csv` (stdin or file) into the few numbers DESIGN.md quotes."""
import csv
import sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_warps',
        'launch__shared_mem_per_block_dynamic', 'launch__shared_mem_per_block_static', 'launch__block_size',
        'launch__grid_size', 'launch__waves_per_multiprocessor',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'smsp__inst_executed.sum', 'lts__t_bytes.sum', 'l1tex__t_bytes_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__t_bytes_pipe_lsu_mem_global_op_st.sum', 'smsp__warps_eligible.avg.per_cycle_active']


def main():
    rows = list(csv.reader(open(sys.argv[1]) if len(sys.argv) > 1 else sys.stdin))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print('----', r[hdr.index('Kernel Name')][:80], 'id', r[0])
        for k in KEYS:
            if k in hdr:
                print(f"  {k:68s} {r[hdr.index(k)]:>18s} {units[hdr.index(k)]}")
        st = []
        for i, h in enumerate(hdr):
            if 'smsp__average_warps_issue_stalled' in h and h.endswith('_per_issue_active.ratio') and r[i] not in ('', 'n/a'):
                st.append((float(r[i].replace(',', '')), h))
        for v, n in sorted(st, reverse=True)[:8]:
            print(f"  stall {v:8.2f} {n.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')}")


if __name__ == '__main__':
    main()
