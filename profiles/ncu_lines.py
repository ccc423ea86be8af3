#!/usr/bin/env python
"""Aggregate an `ncu --page source --csv --print-source cuda,sass` dump by CUDA source line.
usage: ncu -i rep.ncu-rep --page source --csv --print-source cuda,sass --kernel-name regex:<k> > dump.csv
       python profiles/ncu_lines.py dump.csv [top_n]
Prints, for the FIRST kernel instance in the dump, the share of executed warp instructions and of
stall samples per source line (file:line), most expensive first."""
import csv, sys, collections
path = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(open(path, errors='replace')))
inst = collections.Counter(); samp = collections.Counter(); thr = collections.Counter(); src = {}
cur_file = None; hdr = None; kernels = 0; sass_tot = 0
for r in rows:
    if not r: continue
    if r[0] == 'Kernel Name':
        kernels += 1
        if kernels > 1: break
        continue
    if r[0] in ('File Name', 'File Path'): cur_file = r[1].split('/')[-1]; continue
    if r[0] == 'Function Name': continue
    if r[0] == 'Line No': hdr = r; continue
    if hdr is None or len(r) < len(hdr): continue
    if not r[0].strip().isdigit(): continue   # SASS rows repeat what their source line already aggregates
    try:
        ie = float(r[hdr.index('Instructions Executed')] or 0)
        ns = float(r[hdr.index('# Samples')] or 0)
        te = float(r[hdr.index('Thread Instructions Executed')] or 0)
    except ValueError: continue
    key = (cur_file, r[0])
    if r[0]:
        src.setdefault(key, r[1])
    inst[key] += ie; samp[key] += ns; thr[key] += te
ti, ts = sum(inst.values()), sum(samp.values())
print('total warp instructions %.4g, samples %.4g, avg active threads %.1f' % (ti, ts, sum(thr.values()) / max(ti, 1)))
for key, v in inst.most_common(top):
    print('%5.1f%% inst %5.1f%% stall  thr %4.1f  %s:%s  %s' % (100 * v / ti, 100 * samp[key] / max(ts, 1), thr[key] / max(v, 1), key[0], key[1], src.get(key, '')[:90].strip()))
