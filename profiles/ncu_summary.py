#!/usr/bin/env python
"""Summarise an .ncu-rep (ncu --set full) into the handful of metrics DESIGN.md / bench.py quote.
usage: python profiles/ncu_summary.py gpurun_out/prof.ncu-rep [kernel-substring]"""
import csv
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_warps', 'launch__shared_mem_per_block_dynamic',
        'launch__waves_per_multiprocessor', 'launch__grid_size', 'launch__block_size',
        'sm__inst_executed_pipe_fp64', 'sm__pipe_fp64_cycles_active',
        'smsp__inst_executed_pipe_fp64', 'smsp__pipe_fp64_cycles_active',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared', 'smsp__inst_executed.sum',
        'sm__cycles_elapsed.avg ', 'lts__t_sector_hit_rate.pct', 'smsp__issue_active.avg.pct',
        'smsp__average_warp', 'smsp__average_warps_issue_stalled', 'sm__inst_executed_pipe_lsu',
        'l1tex__throughput.avg.pct', 'lts__throughput.avg.pct', 'smsp__thread_inst_executed_per_inst',
        'sm__sass_thread_inst_executed_op_dfma', 'smsp__sass_thread_inst_executed_op_dfma',
        'smsp__sass_thread_inst_executed_op_dadd', 'smsp__sass_thread_inst_executed_op_dmul',
        'smsp__cycles_active.avg', 'sm__cycles_active.avg', 'smsp__warps_eligible',
        'smsp__pcsamp_warps_issue_stalled']


def main():
    rep = sys.argv[1]
    sub = sys.argv[2] if len(sys.argv) > 2 else ''
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True,
                         text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        if sub not in d.get('Kernel Name', ''):
            continue
        print('--- %s  grid %s block %s' % (d['Kernel Name'][:70], d['Grid Size'], d['Block Size']))
        for i, k in enumerate(hdr):
            base = k.split('.', 2)[-1] if k.count('.') >= 2 and k.split('.')[1][0].isupper() else k
            if any(x in k for x in KEYS):
                print('  %-95s %s %s' % (k, d[k], units[i]))
        break


if __name__ == '__main__':
    main()
