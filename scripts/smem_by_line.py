#!/usr/bin/env python
"""Shared-memory wavefronts (total / excessive = bank conflicts) per source line of one profiled launch.
Usage: smem_by_line.py <report.ncu-rep> <launch index> <units> [top]"""
import csv, io, re, subprocess, os, glob, tempfile, sys
rep, launch, units = sys.argv[1], int(sys.argv[2]), float(sys.argv[3])
top = int(sys.argv[4]) if len(sys.argv) > 4 else 30
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
secs = [i for i, r in enumerate(rows) if r and r[0] == 'Kernel Name'] + [len(rows)]
if len(secs) > 2 and all(rows[secs[k]][1] == rows[secs[k + 1]][1] for k in range(0, len(secs) - 1, 2)): launch *= 2
hdr = rows[secs[launch] + 1]; ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[secs[launch] + 2:secs[launch + 1]] if len(r) >= len(hdr)]
kn = rows[secs[launch]][1]
print('kernel:', kn[:70])
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
base = re.sub(r'<.*', '', kn.split('(')[0].split('::')[-1]).split()[-1]
fn = [f for f in per if ('%d%sE' % (len(base), base)) in f or ('%d%sI' % (len(base), base)) in f][0]
lines = per[fn]
if len(lines) != len(data): print('warning: cubin has %d instructions, report %d (different build: line mapping unreliable)' % (len(lines), len(data)))
agg = {}; tw = te = 0
for r, ln in zip(data, lines):
    w = int(r[ix['L1 Wavefronts Shared']] or 0); e = int(r[ix['L1 Wavefronts Shared Excessive']] or 0)
    if not w: continue
    a = agg.setdefault((ln, r[ix['Source']].split()[0 if not r[ix['Source']].startswith('@') else 1]), [0, 0]); a[0] += w; a[1] += e; tw += w; te += e
print('shared wavefronts %.1f per unit, excessive %.1f per unit' % (tw / units, te / units))
srcs = {}
for ((f, n), op), (w, e) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    if f not in srcs:
        p = os.path.join(root, 'arrow-h264_b200', 'csrc', f)
        srcs[f] = open(p).read().splitlines() if os.path.exists(p) else []
    t = srcs[f][n - 1].strip()[:90] if n - 1 < len(srcs[f]) else ''
    print('%7.1f wavefronts/unit %7.1f excessive  %-8s %s:%d  %s' % (w / units, e / units, op, f, n, t))
