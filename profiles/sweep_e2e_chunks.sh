#!/bin/bash
# end-to-end (host buffers) time against the number of batch chunks of the copy/compute pipeline
for c in 4 8 12 16 32; do
  python bench.py --steps 5 --warmup 3 --no-cpu-baseline --e2e-chunks $c 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); e=d['e2e']; print('chunks=$c', 'e2e %.3f ms' % e['ms_per_step'], '%.0f Mpix/s' % e['value'], 'transfers only %.3f ms' % e['transfers_only_ms'])"
done
