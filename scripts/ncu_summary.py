#!/usr/bin/env python
"""Summarises an .ncu-rep (read here, no GPU): per-launch key metrics, and optionally the per-instruction execution
counts of one launch.  Usage: ncu_summary.py <report.ncu-rep> [--sass out.txt [--launch N]]"""
import csv, io, subprocess, sys

KEYS = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__registers_per_thread', 'smsp__inst_executed.sum',
        'sm__inst_executed.avg.per_cycle_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum']


def raw(rep):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[0]
    return hdr, rows[1], rows[2:]


def main():
    rep = sys.argv[1]
    hdr, units, rows = raw(rep)
    ix = {h: i for i, h in enumerate(hdr)}
    stall = [k for k in hdr if 'issue_stalled' in k and 'per_issue_active' in k]
    for r in rows:
        print('---', r[ix['ID']], r[ix['Kernel Name']][:40])
        for k in KEYS:
            if k in ix:
                print('   %-75s %s %s' % (k, r[ix[k]], units[ix[k]]))
        top = sorted(((float(r[ix[k]] or 0), k) for k in stall), reverse=True)[:5]
        print('   stalls per issue: ' + ', '.join('%s %.2f' % (k.split('issue_stalled_')[1].split('_per_')[0], v) for v, k in top))
    if '--sass' in sys.argv:
        outp = sys.argv[sys.argv.index('--sass') + 1]
        launch = int(sys.argv[sys.argv.index('--launch') + 1]) if '--launch' in sys.argv else 0
        out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        secs = [i for i, r in enumerate(rows) if r and r[0] == 'Kernel Name'] + [len(rows)]
        if len(secs) > 2 and all(rows[secs[k]][1] == rows[secs[k + 1]][1] for k in range(0, len(secs) - 1, 2)): launch *= 2   # this ncu prints every launch twice
        hdr = rows[secs[launch] + 1]
        ix = {h: i for i, h in enumerate(hdr)}
        tot = 0
        byop = {}
        with open(outp, 'w') as f:
            for r in rows[secs[launch] + 2:secs[launch + 1]]:
                if len(r) < len(hdr):
                    continue
                n = int(r[ix['Instructions Executed']] or 0)
                tot += n
                src = r[ix['Source']]
                op = (src.split()[1] if src.startswith('@') else src.split()[0]).split('.')[0]
                byop[op] = byop.get(op, 0) + n
                f.write('%s %10d %5s %6s %s\n' % (r[ix['Address']][-5:], n, r[ix['Avg. Threads Executed']][:4], r[ix['# Samples']], src[:100]))
        print('launch %d: %d warp-instructions; by opcode: %s' % (launch, tot, ', '.join('%s %.1f%%' % (k, 100.0 * v / tot) for k, v in sorted(byop.items(), key=lambda kv: -kv[1])[:14])))


if __name__ == '__main__':
    main()
