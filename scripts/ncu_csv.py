#!/usr/bin/env python
"""One CSV row per profiled launch of an .ncu-rep (the summary committed under profiles/).  Usage: ncu_csv.py <rep> > out.csv"""
import csv, io, subprocess, sys
KEYS = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread', 'smsp__inst_executed.sum',
        'sm__inst_executed.avg.per_cycle_active', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'dram__bytes_read.sum',
        'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct',
        'l1tex__t_sector_hit_rate.pct', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'lts__t_bytes.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum']
out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units, data = rows[0], rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
keys = [k for k in KEYS if k in ix]
stall = [k for k in hdr if 'issue_stalled' in k and 'per_issue_active' in k]
w = csv.writer(sys.stdout)
w.writerow(['ID', 'Kernel Name'] + keys + ['top stalls (cycles per issue)'])
w.writerow(['', ''] + [units[ix[k]] for k in keys] + [''])
for r in data:
    top = sorted(((float(r[ix[k]] or 0), k) for k in stall), reverse=True)[:4]
    w.writerow([r[ix['ID']], r[ix['Kernel Name']]] + [r[ix[k]] for k in keys] +
               ['; '.join('%s %.2f' % (k.split('issue_stalled_')[1].split('_per_')[0], v) for v, k in top)])
