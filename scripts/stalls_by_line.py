#!/usr/bin/env python
"""Warp-stall samples and executed instructions per source line of one profiled launch (ncu source page joined with
the cubin's line table).  Usage: stalls_by_line.py <report.ncu-rep> <launch index> <kernel substring> <units> [top]"""
import csv, io, re, subprocess, os, glob, tempfile, sys
rep, launch, kname, units = sys.argv[1], int(sys.argv[2]), sys.argv[3], float(sys.argv[4])
top = int(sys.argv[5]) if len(sys.argv) > 5 else 40
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
secs = [i for i, r in enumerate(rows) if r and r[0] == 'Kernel Name'] + [len(rows)]
if len(secs) > 2 and all(rows[secs[k]][1] == rows[secs[k + 1]][1] for k in range(0, len(secs) - 1, 2)): launch *= 2   # this ncu prints every launch twice
hdr = rows[secs[launch] + 1]; ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[secs[launch] + 2:secs[launch + 1]] if len(r) >= len(hdr)]
tmp = tempfile.mkdtemp()
subprocess.run(['cuobjdump', '-xelf', 'all', os.path.join(root, 'arrow-h264_b200', 'libh264recon.so')], cwd=tmp, capture_output=True)
cubin = [f for f in glob.glob(os.path.join(tmp, '*.cubin')) if os.path.basename(f).startswith('kernels')][0]
dis = subprocess.run(['nvdisasm', '-g', '-c', cubin], capture_output=True, text=True).stdout
cur = None; line = None; per = {}
for l in dis.splitlines():
    m = re.match(r'\s*\.text\.(\S+):', l)
    if m: cur = m.group(1); per[cur] = []; continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: line = (os.path.basename(m.group(1)), int(m.group(2))); continue
    if cur and re.match(r'\s+/\*[0-9a-f]{4,}\*/', l): per[cur].append(line)
kn = rows[secs[launch]][1]
print('section kernel:', kn[:60])
base = kn.split('(')[0].split('::')[-1]
fn = [f for f in per if ('%d%sE' % (len(base), base)) in f][0]
lines = per[fn]
if len(lines) != len(data): print('warning: cubin has %d instructions, report %d (different build)' % (len(lines), len(data)))
agg = {}; tot = 0; tote = 0
for r, ln in zip(data, lines):
    n = int(r[ix['# Samples']] or 0); e = int(r[ix['Instructions Executed']] or 0)
    a = agg.setdefault(ln, [0, 0]); a[0] += n; a[1] += e; tot += n; tote += e
srcs = {}
print('total samples %d, executed warp-instructions %d (%.1f per unit)' % (tot, tote, tote / units))
for (f, n), (s, e) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    if f not in srcs:
        p = os.path.join(root, 'arrow-h264_b200', 'csrc', f)
        srcs[f] = open(p).read().splitlines() if os.path.exists(p) else []
    t = srcs[f][n - 1].strip()[:100] if n - 1 < len(srcs[f]) else ''
    print('%5.1f%% samples  %8.1f instr/unit  %s:%d  %s' % (100.0 * s / max(tot, 1), e / units, f, n, t))
