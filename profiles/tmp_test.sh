for v in "" "-DDIRT_BWD_WARPS_C4=1"; do
  DIRT_NVCC_EXTRA="$v" python -c "from dirt_b200 import build; build.build(force=True)"
  DIRT_NVCC_EXTRA="$v" python -m pytest tests -m gpu -x -q 2>&1 | tail -2
  for wl in cfg3 cfg3 cfg5 cfg4; do
  DIRT_NVCC_EXTRA="$v" python bench.py --workload $wl --steps 20 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('[$v] $wl', 'step %.4f ms' % d['ms_per_step'], 'fwd_k %.4f' % d['roofline']['forward_kernel']['ms'], 'bwd_k %.4f' % d['roofline']['backward_kernel']['ms'], 'Mpix/s %.0f' % d['value'])"
  done
done
