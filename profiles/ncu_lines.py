#!/usr/bin/env python
"""Aggregate an `ncu --page source --print-source cuda,sass --csv` dump by source line.
usage: ncu_lines.py file.csv [top_n | all]   (all: every line of backward.cu / raster.cu / common.cuh in file order)"""
import csv, sys, collections
path = sys.argv[1]; every = len(sys.argv) > 2 and sys.argv[2] == 'all'; top = int(sys.argv[2]) if len(sys.argv) > 2 and not every else 40
cur_file = None; hdr = None
agg = collections.OrderedDict()
for r in csv.reader(open(path)):
    if not r: continue
    if r[0] == 'File Path': cur_file = r[1].split('/')[-1]; continue
    if r[0] == 'Function Name': continue
    if r[0] == 'Line No': hdr = r; continue
    if hdr is None or len(r) < len(hdr): continue
    if r[2] != '-': continue          # SASS rows carry an address; source rows have '-' and hold the line totals
    d = dict(zip(hdr, r))
    try:
        inst = int(d['Instructions Executed']); samp = int(d['# Samples'])
    except ValueError:
        continue
    key = (cur_file, int(r[0]))
    if key in agg:
        agg[key][0] += inst; agg[key][1] += samp
    else:
        agg[key] = [inst, samp, r[1].strip()[:110]]
tot_i = sum(v[0] for v in agg.values()) or 1; tot_s = sum(v[1] for v in agg.values()) or 1
print('total inst %d  samples %d' % (tot_i, tot_s))
if every:
    for k, v in sorted(agg.items()):
        if v[0] or v[1]:
            print('%-14s %5d %6.2f%% inst %6.2f%% smp  %s' % (k[0], k[1], 100.0 * v[0] / tot_i, 100.0 * v[1] / tot_s, v[2][:90]))
    sys.exit(0)
print('--- by instructions')
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print('%5.1f%% inst %5.1f%% smp  %s:%d  %s' % (100.0 * v[0] / tot_i, 100.0 * v[1] / tot_s, k[0], k[1], v[2]))
print('--- by stall samples')
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print('%5.1f%% smp %5.1f%% inst  %s:%d  %s' % (100.0 * v[1] / tot_s, 100.0 * v[0] / tot_i, k[0], k[1], v[2]))
