#!/bin/bash
# direct-add criterion: lanes holding a record of the face (one vote) instead of the record count (four votes)
for t in 0 3 4 5 6 8; do
  DIRT_NVCC_EXTRA="-DDIRT_BWD_SMALL_LANES=$t" python -c "from dirt_b200 import build; build.build(force=True)"
  for wl in cfg3 cfg5; do
    python bench.py --workload $wl --steps 20 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('lanes=$t $wl', 'step %.4f ms' % d['ms_per_step'], 'fwd_k %.4f' % d['roofline']['forward_kernel']['ms'], 'bwd_k %.4f' % d['roofline']['backward_kernel']['ms'], 'Mpix/s %.0f' % d['value'])"
  done
done
python -c "from dirt_b200 import build; build.build(force=True)"
