#!/bin/bash
# forward: warp-phase loop trimmed (113 -> 87 SASS instructions); backward: larger direct-add thresholds
for t in 12 16 24 32; do
  DIRT_NVCC_EXTRA="-DDIRT_BWD_SMALL_FACE=$t" python -c "from dirt_b200 import build; build.build(force=True)"
  for wl in cfg3 cfg5; do
    python bench.py --workload $wl --steps 20 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('small=$t $wl', 'step %.4f ms' % d['ms_per_step'], 'fwd_k %.4f' % d['roofline']['forward_kernel']['ms'], 'bwd_k %.4f' % d['roofline']['backward_kernel']['ms'], 'Mpix/s %.0f' % d['value'])"
  done
done
python -c "from dirt_b200 import build; build.build(force=True)"
