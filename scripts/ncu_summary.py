#!/usr/bin/env python
"""Summarise an .ncu-rep (raw page + SASS regions) into text.  Usage: ncu_summary.py rep [out.txt]"""
import csv, subprocess, sys, io
rep = sys.argv[1]
out = open(sys.argv[2], 'w') if len(sys.argv) > 2 else sys.stdout
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
keep = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'launch__shared_mem_per_block_dynamic',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio']
for vals in rows[2:]:
    for i, h in enumerate(hdr):
        if h in keep:
            out.write('%-82s %-14s %s\n' % (h, units[i], vals[i]))
    out.write('\n')
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == 'Address')
hdr = rows[hi]; data = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
ix = {h: i for i, h in enumerate(hdr)}
tot = sum(float(r[ix['Instructions Executed']]) for r in data)
tthr = sum(float(r[ix['Thread Instructions Executed']]) for r in data)
out.write('SASS regions (contiguous instructions with equal execution count), share of warp instructions:\n')
out.write('total warp inst %.4e  thread inst %.4e  avg active threads %.2f  SASS lines %d\n' % (tot, tthr, tthr / tot, len(data)))
seg, cur = [], None
for k, r in enumerate(data):
    ex = float(r[ix['Instructions Executed']]); th = float(r[ix['Avg. Threads Executed']])
    if cur and abs(cur['ex'] - ex) <= 0.03 * max(cur['ex'], ex) + 1e4:
        cur['n'] += 1; cur['sum'] += ex; cur['thr'] += ex * th; cur['end'] = k
    else:
        cur = dict(start=k, end=k, ex=ex, n=1, sum=ex, thr=ex * th, first=r[ix['Source']].strip()); seg.append(cur)
for s in seg:
    if s['sum'] / tot > 0.004:
        out.write('%5d-%5d n=%4d exec/inst=%9.3fM share=%5.1f%% avgthr=%5.1f  %s\n' % (
            s['start'], s['end'], s['n'], s['ex'] / 1e6, 100 * s['sum'] / tot, s['thr'] / s['sum'], s['first'][:70]))
