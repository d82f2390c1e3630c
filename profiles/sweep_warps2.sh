#!/bin/bash
# one or two warps per CTA (32 / 16 CTAs per SM): cheaper index arithmetic, independent retirement
for cfg in "4 8 4 8" "1 32 4 8" "4 8 1 32" "2 16 2 16" "1 32 1 32"; do
  set -- $cfg
  DIRT_NVCC_EXTRA="-DDIRT_RASTER_WARPS=$1 -DDIRT_RASTER_MIN_BLOCKS=$2 -DDIRT_BWD_WARPS=$3 -DDIRT_BWD_MIN_BLOCKS=$4" python -c "from dirt_b200 import build; build.build(force=True)"
  for wl in cfg3 cfg5; do
  python bench.py --workload $wl --steps 20 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('raster warps=$1 min=$2  bwd warps=$3 min=$4 $wl', 'step %.4f ms' % d['ms_per_step'], 'fwd_k %.4f' % d['roofline']['forward_kernel']['ms'], 'bwd_k %.4f' % d['roofline']['backward_kernel']['ms'])"
  done
done
python -c "from dirt_b200 import build; build.build(force=True)"
