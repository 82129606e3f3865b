#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel: launches, total time, share, and the
grid size (CTAs) with the CTA-time product sum(grid * time) -- a proxy for how much of the machine a kernel holds when many
instances share the GPU.
    python tools/summarise_launches.py gpurun_out/launches.csv "comment line" > profiles/<name>_summary.csv"""
import csv
import re
import sys
from collections import defaultdict

rows = []
with open(sys.argv[1], newline='') as f:
    lines = [l for l in f if not l.startswith('==')]
rd = csv.reader(lines)
hdr = next(rd)
ik, iv, iu = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
ig, im = hdr.index('Grid Size'), hdr.index('Metric Name')
agg = defaultdict(list)
ctas = defaultdict(list)
for r in rd:
    if len(r) <= iv or r[im] != 'gpu__time_duration.sum':
        continue
    name = re.sub(r'^void ', '', r[ik])
    name = re.sub(r'\(.*$', '', name).replace('<unnamed>::', '').replace('(anonymous namespace)::', '')
    v = float(r[iv].replace(',', ''))
    v *= {'ns': 1e-3, 'us': 1.0, 'usecond': 1.0, 'ms': 1e3, 'msecond': 1e3, 'nsecond': 1e-3}.get(r[iu], 1.0)
    agg[name].append(v)
    g = 1
    for x in re.findall(r'\d+', r[ig]):
        g *= int(x)
    ctas[name].append((g, v))
tot = sum(sum(v) for v in agg.values())
n = sum(len(v) for v in agg.values())
print('# %s; %d launches' % (sys.argv[2] if len(sys.argv) > 2 else sys.argv[1], n))
print('# per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes')
ctot = sum(g * v for l in ctas.values() for g, v in l)
print('kernel,launches,sum_us,share_pct,median_us,max_us,median_ctas,cta_time_share_pct')
for name, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    v = sorted(v)
    gs = sorted(g for g, _ in ctas[name])
    print('"%s",%d,%.1f,%.2f,%.2f,%.2f,%d,%.2f' % (name, len(v), sum(v), 100 * sum(v) / tot, v[len(v) // 2], v[-1], gs[len(gs) // 2],
                                                100 * sum(g * t for g, t in ctas[name]) / ctot))
