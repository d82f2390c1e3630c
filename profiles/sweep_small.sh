#!/bin/bash
# threshold (records per tile) below which a face's terms bypass the butterfly and are added directly
for t in 0 2 4 8 12; do
  DIRT_NVCC_EXTRA="-DDIRT_BWD_SMALL_FACE=$t" python -c "from dirt_b200 import build; build.build(force=True)"
  for wl in cfg3 cfg5; do
    python bench.py --workload $wl --steps 10 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('small=$t $wl', 'step %.3f ms' % d['ms_per_step'], 'bwd_k %.3f' % d['roofline']['backward_kernel']['ms'], 'Mpix/s %.0f' % d['value'])"
  done
done
python -c "from dirt_b200 import build; build.build(force=True)"
