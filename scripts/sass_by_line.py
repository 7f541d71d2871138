#!/usr/bin/env python
"""Per-source-line executed warp-instruction counts of one profiled launch: joins the per-instruction counters of an
.ncu-rep (source page, SASS view) with the line table of the cubin (nvdisasm -g), instruction by instruction.
Usage: sass_by_line.py <report.ncu-rep> <launch index> <kernel name substring> [--per N] [--top K]"""
import csv, io, re, subprocess, sys, tempfile, os, glob

rep, launch, kname = sys.argv[1], int(sys.argv[2]), sys.argv[3]
per = float(sys.argv[sys.argv.index('--per') + 1]) if '--per' in sys.argv else 1.0
top = int(sys.argv[sys.argv.index('--top') + 1]) if '--top' in sys.argv else 60
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(root, 'arrow-h264_b200', 'libh264recon.so')

out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
secs = [i for i, r in enumerate(rows) if r and r[0] == 'Kernel Name'] + [len(rows)]
if len(secs) > 2 and all(rows[secs[k]][1] == rows[secs[k + 1]][1] for k in range(0, len(secs) - 1, 2)): launch *= 2   # this ncu prints every launch twice
hdr = rows[secs[launch] + 1]
ix = {h: i for i, h in enumerate(hdr)}
counts = [(r[ix['Source']], int(r[ix['Instructions Executed']] or 0)) for r in rows[secs[launch] + 2:secs[launch + 1]] if len(r) >= len(hdr)]

tmp = tempfile.mkdtemp()
subprocess.run(['cuobjdump', '-xelf', 'all', so], cwd=tmp, capture_output=True)
cubin = [f for f in glob.glob(os.path.join(tmp, '*.cubin')) if os.path.basename(f).startswith('kernels')][0]
dis = subprocess.run(['nvdisasm', '-g', '-c', cubin], capture_output=True, text=True).stdout
cur_fn, line, per_fn = None, None, {}
for l in dis.splitlines():
    m = re.match(r'\s*\.text\.(\S+):', l)
    if m:
        cur_fn = m.group(1); per_fn[cur_fn] = []; continue
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*inlined at "([^"]+)", line (\d+))?', l)
    if m:
        line = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if cur_fn and re.match(r'\s+/\*[0-9a-f]{4,}\*/', l):
        per_fn[cur_fn].append(line)
fn = [f for f in per_fn if kname in f][0]
lines = per_fn[fn]
if len(lines) != len(counts):
    print('warning: %d SASS instructions in the cubin vs %d in the report (different build?)' % (len(lines), len(counts)))
agg = {}
for (src, n), ln in zip(counts, lines):
    agg[ln] = agg.get(ln, 0) + n
tot = sum(n for _, n in counts)
print('%s launch %d: %d warp-instructions (%.1f per unit)' % (fn[:40], launch, tot, tot / per))
srcs = {}
for (f, n), v in sorted(agg.items(), key=lambda kv: -kv[1])[:top]:
    if f not in srcs:
        p = os.path.join(root, 'arrow-h264_b200', 'csrc', f)
        srcs[f] = open(p).read().splitlines() if os.path.exists(p) else []
    text = srcs[f][n - 1].strip()[:90] if n - 1 < len(srcs[f]) else ''
    print('%6.1f  %5.1f%%  %s:%d  %s' % (v / per, 100.0 * v / tot, f, n, text))
