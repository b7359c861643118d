#!/usr/bin/env python
"""Prints the key ncu metrics of every kernel in a .ncu-rep (raw page), for profiles/."""
import csv, subprocess, sys, io
KEYS = ['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
 'launch__registers_per_thread','launch__grid_size','launch__occupancy_limit_registers','sm__warps_active.avg.pct_of_peak_sustained_active',
 'smsp__inst_executed.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
 'l1tex__t_sector_hit_rate.pct','lts__t_sector_hit_rate.pct','sm__cycles_elapsed.max','smsp__cycles_active.avg',
 'smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio','smsp__average_warp_latency_issue_stalled_short_scoreboard.ratio',
 'smsp__average_warp_latency_issue_stalled_wait.ratio','smsp__average_warp_latency_issue_stalled_math_pipe_throttle.ratio',
 'smsp__average_warp_latency_issue_stalled_not_selected.ratio','smsp__average_warp_latency_issue_stalled_dispatch_stall.ratio',
 'smsp__average_warp_latency_issue_stalled_barrier.ratio','smsp__average_warp_latency_issue_stalled_lg_throttle.ratio',
 'smsp__average_warp_latency_issue_stalled_mio_throttle.ratio','smsp__average_warp_latency_issue_stalled_branch_resolving.ratio',
 'smsp__average_warp_latency_issue_stalled_no_instruction.ratio','smsp__average_warp_latency_issue_stalled_imc_miss.ratio',
 'smsp__warps_eligible.avg.per_cycle_active','smsp__warps_active.avg.per_cycle_active','sm__warps_active.avg.per_cycle_active']
for rep in sys.argv[1:]:
    out = subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print('==', rep, '|', d.get('Kernel Name'))
        for k in KEYS:
            if k in d: print('  %-75s %s %s' % (k, d[k], units[hdr.index(k)]))
