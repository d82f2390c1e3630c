#!/bin/bash
# L2 prefetch distance (thread blocks ahead in launch order) and image order of the backward kernel
for v in "" "-DDIRT_BWD_PREFETCH_BLOCKS=256" "-DDIRT_BWD_PREFETCH_BLOCKS=1024" "-DDIRT_BWD_PREFETCH_BLOCKS=4096" \
         "-DDIRT_BWD_PREFETCH_BLOCKS=16384" "-DDIRT_BWD_REVERSE=1" "-DDIRT_BWD_REVERSE=1 -DDIRT_BWD_PREFETCH_BLOCKS=1024"; do
  DIRT_NVCC_EXTRA="$v" python -c "from dirt_b200 import build; build.build(force=True)"
  for wl in cfg3 cfg5; do
    python bench.py --workload $wl --steps 20 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('[$v] $wl', 'step %.4f ms' % d['ms_per_step'], 'fwd_k %.4f' % d['roofline']['forward_kernel']['ms'], 'bwd_k %.4f' % d['roofline']['backward_kernel']['ms'], 'Mpix/s %.0f' % d['value'])"
  done
done
python -c "from dirt_b200 import build; build.build(force=True)"
