#!/usr/bin/env python
"""profiles/traffic.json (DRAM bytes per launch of the two hot kernels, read by bench.py for roofline.traffic) from an
`ncu -i x.ncu-rep --page raw --csv` dump.  usage: make_traffic.py raw.csv workload summary_name [images_per_launch]"""
import csv, json, os, sys
raw, workload, summary = sys.argv[1], sys.argv[2], sys.argv[3]
images = int(sys.argv[4]) if len(sys.argv) > 4 else None   # batch size of the captured launch (None = the workload's own)
rows = list(csv.reader(open(raw)))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
def to_bytes(v, u):
    v = float(v.replace(',', ''))
    return v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}[u]
out = {}
for r in rows[2:]:
    name = r[idx['Kernel Name']]
    key = 'forward' if 'raster_kernel' in name else 'backward' if 'backward' in name else None
    if key is None or key in out:
        continue
    rd = to_bytes(r[idx['dram__bytes_read.sum']], units[idx['dram__bytes_read.sum']])
    wr = to_bytes(r[idx['dram__bytes_write.sum']], units[idx['dram__bytes_write.sum']])
    out[key] = {'dram_bytes': int(rd + wr), 'images': images,
                'source': '%s: %s dram__bytes_read.sum %.1f MB + dram__bytes_write.sum %.1f MB per launch' % (summary, name.split('(')[0], rd / 1e6, wr / 1e6)}
path = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'traffic.json')
allw = json.load(open(path)) if os.path.exists(path) else {}
allw[workload] = out
json.dump(allw, open(path, 'w'), indent=1)
print(json.dumps(out, indent=1))
