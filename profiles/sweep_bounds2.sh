#!/bin/bash
for cfg in "8 8" "7 7" "8 7" "7 8"; do
  set -- $cfg
  DIRT_NVCC_EXTRA="-DDIRT_RASTER_MIN_BLOCKS=$1 -DDIRT_BWD_MIN_BLOCKS=$2" python -c "from dirt_b200 import build; build.build(force=True)"
  python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('raster_min=$1 bwd_min=$2', 'step %.3f ms' % d['ms_per_step'], 'fwd_k %.3f' % d['roofline']['forward_kernel']['ms'], 'bwd_k %.3f' % d['roofline']['backward_kernel']['ms'])"
done
python -c "from dirt_b200 import build; build.build(force=True)"
