#!/usr/bin/env python
"""Opcode histogram (weighted by executions) from an `ncu --page source --print-source cuda,sass --csv` dump."""
import csv, sys, collections
hdr = None; hist = collections.Counter(); tot = 0
for r in csv.reader(open(sys.argv[1])):
    if not r: continue
    if r[0] == 'Line No': hdr = r; continue
    if hdr is None or len(r) < len(hdr) or r[2] == '-': continue
    d = dict(zip(hdr, r))
    try: n = int(d['Instructions Executed'])
    except ValueError: continue
    sass = r[3].strip()
    toks = sass.split()
    op = toks[1] if toks and toks[0].startswith('@') and len(toks) > 1 else (toks[0] if toks else '?')
    op = op.split('.')[0] + ('.' + op.split('.')[1] if op.startswith(('LD', 'ST', 'ATOM', 'RED', 'SHFL', 'MUFU', 'LDG', 'STG')) and '.' in op else '')
    hist[op] += n; tot += n
print('total', tot)
for op, n in hist.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 40):
    print('%6.2f%%  %12d  %s' % (100.0 * n / tot, n, op))
