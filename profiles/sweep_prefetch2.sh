#!/bin/bash
# same-tile L2 prefetch issued before the tile's first dependent load returns (forward: background, backward: grad_pixels)
for v in "" "-DDIRT_RASTER_PREFETCH_BG=1" "-DDIRT_BWD_PREFETCH_GP=1"; do
  DIRT_NVCC_EXTRA="$v" python -c "from dirt_b200 import build; build.build(force=True)"
  for wl in cfg3 cfg5; do
  python bench.py --workload $wl --steps 20 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('[$v] $wl', 'step %.4f ms' % d['ms_per_step'], 'fwd_k %.4f' % d['roofline']['forward_kernel']['ms'], 'bwd_k %.4f' % d['roofline']['backward_kernel']['ms'])"
  done
done
python -c "from dirt_b200 import build; build.build(force=True)"
