#!/usr/bin/env python
"""Summarise an .ncu-rep: key metrics (raw page) and the hottest source regions (source page).
usage: ncu_summary.py report.ncu-rep [--hot N]"""
import csv, subprocess, sys, io
rep = sys.argv[1]
KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct',
        'smsp__issue_active.avg.pct', 'sm__inst_executed_pipe_fma.avg.pct', 'smsp__inst_executed.sum ',
        'sm__warps_active.avg.pct', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread ',
        'launch__shared_mem_per_block_dynamic', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct',
        'smsp__average_warps_issue_stalled', 'l1tex__t_sector_hit_rate', 'lts__t_sector_hit_rate.pct',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'sm__throughput.avg.pct', 'l1tex__throughput.avg.pct',
        'lts__throughput.avg.pct', 'sm__pipe_tensor', 'sm__inst_executed_pipe_alu.avg.pct', 'sm__inst_executed_pipe_lsu.avg.pct',
        'sm__inst_executed_pipe_xu.avg.pct', 'smsp__inst_executed_op_shared', 'sm__inst_executed_pipe_fmaheavy', 'sm__inst_executed_pipe_fmalite']
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
r = list(csv.reader(io.StringIO(raw)))
hdr, units = r[0], r[1]
for row in r[2:]:
    print('== kernel:', row[hdr.index('Kernel Name')][:100])
    for h, u, v in zip(hdr, units, row):
        if any(k in h + ' ' for k in KEYS):
            try:
                if float(v) == 0: continue
            except ValueError: pass
            print(f'  {h:88s} {u:14s} {v}')
if '--hot' in sys.argv:
    n = int(sys.argv[sys.argv.index('--hot') + 1])
    src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
    r = list(csv.reader(io.StringIO(src)))
    hdr = r[1]; rows = r[2:]
    isrc, isamp, iex = hdr.index('Source'), hdr.index('# Samples'), hdr.index('Instructions Executed')
    tot = sum(int(x[isamp]) for x in rows)
    print('total samples', tot, 'sass instructions', len(rows))
    top = sorted(range(len(rows)), key=lambda i: -int(rows[i][isamp]))[:n]
    for i in sorted(top):
        print(f'  {i:5d} {100*int(rows[i][isamp])/tot:5.2f}% x{rows[i][iex]:>12s}  {rows[i][isrc][:90]}')
