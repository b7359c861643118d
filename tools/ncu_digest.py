#!/usr/bin/env python
"""Digest of an ncu report: per kernel the headline metrics, stall reasons per issued instruction, and (with --hot)
the instruction mix of the hottest loop body (instructions sharing the modal execution count).

    python tools/ncu_digest.py gpurun_out/x.ncu-rep [--hot]"""
import collections
import csv
import io
import re
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}
want = ['gpu__time_duration.sum', 'launch__registers_per_thread', 'sm__warps_active.avg.per_cycle_active', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum']
stall = [h for h in hdr if 'smsp__average_warps_issue_stalled' in h and 'per_issue_active' in h]
seen = set()
for n, r in enumerate(data):
    name = r[idx['Kernel Name']]
    if name in seen:
        continue
    seen.add(name)
    print('==', name[:70])
    for w in want:
        if w in idx:
            print('   %-72s %s %s' % (w, r[idx[w]], units[idx[w]]))
    st = sorted(((float(r[idx[h]].replace(',', '')), h) for h in stall if r[idx[h]] not in ('', 'n/a')), reverse=True)
    print('   stalls per issued instruction: ' + ', '.join('%s %.2f' % (h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''), v) for v, h in st[:9]))
    if '--hot' in sys.argv:
        src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", str(n), "--launch-count", "1"], capture_output=True, text=True).stdout
        srows = list(csv.reader(io.StringIO(src)))
        sh = srows[1]
        si = {h: i for i, h in enumerate(sh)}
        ins, last = [], None
        for q in srows[2:]:
            if len(q) < len(sh) or not q[si['Instructions Executed']].isdigit():
                continue
            key = q[si['Address']]
            if key == last:
                continue                      # (the csv lists every instruction twice)
            last = key
            ins.append((q[si['Source']], int(q[si['Instructions Executed']]), int(q[si['# Samples']] or 0)))
        tot = sum(n_ for _, n_, _ in ins)
        cnt = collections.Counter()
        for _, n_, _ in ins:
            cnt[n_] += n_
        mode, share = cnt.most_common(1)[0]
        hot = [(s_, n_, smp) for s_, n_, smp in ins if abs(n_ - mode) <= 0.06 * mode]
        op = collections.Counter()
        for s_, n_, _ in hot:
            t = re.sub(r'^@!?U?P\w+\s+', '', s_.strip())
            op[t.split()[0].split('.')[0]] += 1
        print('   executed %d instructions; hot body: %d static instructions x %d executions = %.0f %% of all' % (tot, len(hot), mode, 100.0 * sum(n_ for _, n_, _ in hot) / tot))
        print('   hot body mix: ' + ' '.join('%s=%d' % kv for kv in op.most_common(30)))
