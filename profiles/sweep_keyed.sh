#!/bin/bash
# on-box comparison: per-face transposed butterfly vs keyed segmented reduction, cfg3 / cfg4 / cfg5
for keyed in 0 1; do
  DIRT_NVCC_EXTRA="-DDIRT_BWD_KEYED=$keyed" python -c "from dirt_b200 import build; build.build(force=True)"
  for wl in cfg3 cfg4 cfg5; do
    python bench.py --workload $wl --steps 10 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('keyed=$keyed $wl', 'step %.3f ms' % d['ms_per_step'], 'fwd_k %.3f' % d['roofline']['forward_kernel']['ms'], 'bwd_k %.3f' % d['roofline']['backward_kernel']['ms'], 'Mpix/s %.0f' % d['value'], 'frac %.3f' % d['roofline']['step']['frac'])"
  done
done
python -c "from dirt_b200 import build; build.build(force=True)"
