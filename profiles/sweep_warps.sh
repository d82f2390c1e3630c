#!/bin/bash
# warps per CTA (min blocks per SM scaled to keep 32 warps/SM resident)
for cfg in "4 8 4 8" "2 16 2 16" "8 4 8 4" "2 16 4 8" "8 4 4 8"; do
  set -- $cfg
  DIRT_NVCC_EXTRA="-DDIRT_RASTER_WARPS=$1 -DDIRT_RASTER_MIN_BLOCKS=$2 -DDIRT_BWD_WARPS=$3 -DDIRT_BWD_MIN_BLOCKS=$4" python -c "from dirt_b200 import build; build.build(force=True)"
  python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('raster warps=$1 min=$2  bwd warps=$3 min=$4', 'step %.3f ms' % d['ms_per_step'], 'fwd_k %.3f' % d['roofline']['forward_kernel']['ms'], 'bwd_k %.3f' % d['roofline']['backward_kernel']['ms'])"
done
python -c "from dirt_b200 import build; build.build(force=True)"
