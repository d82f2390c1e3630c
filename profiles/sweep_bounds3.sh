#!/bin/bash
# more CTAs per SM (fewer registers) than the 8 x 64-register default
for cfg in "8 8" "9 8" "10 8" "12 8" "8 9" "8 10"; do
  set -- $cfg
  DIRT_NVCC_EXTRA="-DDIRT_RASTER_MIN_BLOCKS=$1 -DDIRT_BWD_MIN_BLOCKS=$2" python -c "from dirt_b200 import build; build.build(force=True)"
  for wl in cfg3 cfg5; do
  python bench.py --workload $wl --steps 20 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('raster_min=$1 bwd_min=$2 $wl', 'step %.4f ms' % d['ms_per_step'], 'fwd_k %.4f' % d['roofline']['forward_kernel']['ms'], 'bwd_k %.4f' % d['roofline']['backward_kernel']['ms'])"
  done
done
python -c "from dirt_b200 import build; build.build(force=True)"
