#!/bin/bash
# setup kernel: threads per block / register cap (phases_ms.forward_call - forward kernel = memset + setup + scan + fill)
for v in "" "-DDIRT_SETUP_MIN_BLOCKS=4" "-DDIRT_SETUP_THREADS=128 -DDIRT_SETUP_MIN_BLOCKS=8" "-DDIRT_SETUP_THREADS=128" "-DDIRT_SETUP_THREADS=64" "-DDIRT_SETUP_THREADS=128 -DDIRT_SETUP_MIN_BLOCKS=10"; do
  DIRT_NVCC_EXTRA="$v" python -c "from dirt_b200 import build; build.build(force=True)"
  for wl in cfg3 cfg5; do
    python bench.py --workload $wl --steps 20 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('[$v] $wl', 'step %.4f ms' % d['ms_per_step'], 'fwd_call %.4f' % d['phases_ms']['forward_call'], 'fwd_k %.4f' % d['roofline']['forward_kernel']['ms'], 'setup etc %.4f' % (d['phases_ms']['forward_call'] - d['roofline']['forward_kernel']['ms']))"
  done
done
python -c "from dirt_b200 import build; build.build(force=True)"
