#!/usr/bin/env python
"""
Summarise Nsight Compute output into the small text files kept under profiles/.

  python tools/ncu_summary.py launches gpurun_out/launches.csv > profiles/rNN_launches.md
  python tools/ncu_summary.py kernel gpurun_out/prof.ncu-rep   > profiles/rNN_kernel.md
"""
import collections
import csv
import subprocess
import sys

KEEP = [
    'gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size',
    'launch__registers_per_thread', 'launch__occupancy_limit_registers',
    'launch__occupancy_limit_shared_mem', 'launch__shared_mem_per_block_static',
    'sm__warps_active.avg.pct_of_peak_sustained_active',
    'sm__throughput.avg.pct_of_peak_sustained_elapsed',
    'smsp__issue_active.avg.pct_of_peak_sustained_active',
    'smsp__inst_executed.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio',
    'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
    'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
    'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
    'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
    'lts__throughput.avg.pct_of_peak_sustained_elapsed',
    'lts__t_sector_hit_rate.pct', 'lts__t_sectors.sum',
    'lts__t_requests_srcunit_tex_op_red.sum', 'lts__t_requests_srcunit_tex_op_read.sum',
    'dram__bytes_read.sum', 'dram__bytes_write.sum',
    'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
    'sm__cycles_elapsed.avg', 'sm__cycles_elapsed.avg.per_second',
    'smsp__warps_eligible.avg.per_cycle_active',
    'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio',
    'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
    'smsp__inst_executed_op_shared_ld.sum', 'smsp__inst_executed_op_shared_st.sum',
    'smsp__inst_executed_op_global_red.sum', 'smsp__inst_executed_op_global_ld.sum',
    'sm__sass_inst_executed_op_global_red.sum',
]


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10 and r[0].isdigit()]
    names = [r[4].split('(')[0].replace('void ', '') for r in rows]
    t = [float(r[-1]) for r in rows]
    total = sum(t)
    agg = collections.OrderedDict()
    for n, v in zip(names, t):
        a = agg.setdefault(n, [0.0, 0])
        a[0] += v
        a[1] += 1
    print('| kernel | launches | total us | mean us | share |')
    print('|---|---:|---:|---:|---:|')
    for n, (v, c) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        print('| %s | %d | %.1f | %.1f | %.1f%% |' % (n, c, v / 1e3, v / c / 1e3, 100 * v / total))
    print('\n%d launches, %.3f ms of kernel time' % (len(rows), total / 1e6))


def kernel(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'],
                         stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        d = dict(zip(hdr, zip(units, vals)))
        print('## %s  grid %s block %s\n' % (d.get('Kernel Name', ('', '?'))[1],
                                              d.get('Grid Size', ('', '?'))[1],
                                              d.get('Block Size', ('', '?'))[1]))
        print('| metric | value | unit |')
        print('|---|---:|---|')
        for k in KEEP:
            if k in d:
                print('| %s | %s | %s |' % (k, d[k][1], d[k][0]))
        print()


if __name__ == '__main__':
    {'launches': launches, 'kernel': kernel}[sys.argv[1]](sys.argv[2])
