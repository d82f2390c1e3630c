#!/bin/bash
# timing-only ablations of the backward kernel (results are wrong when DIRT_ABLATE != 0)
for a in 0 1 2 3; do
  DIRT_NVCC_EXTRA="-DDIRT_ABLATE=$a" python -c "from dirt_b200 import build; build.build(force=True)"
  python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('ablate=$a (1: no face loop, 2: no Scharr/dilation/position terms, 3: both)', 'bwd_k %.3f ms' % d['roofline']['backward_kernel']['ms'])"
done
python -c "from dirt_b200 import build; build.build(force=True)"
