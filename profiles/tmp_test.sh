DIRT_NVCC_EXTRA="-DDIRT_BWD_WARPS_C4=1" python -c "from dirt_b200 import build; build.build(force=True)"
DIRT_NVCC_EXTRA="-DDIRT_BWD_WARPS_C4=1" python -m pytest tests -m gpu -q -k "random_soup or reference_scenes or reduced or edge_cases or large_faces" 2>&1 | grep -E "FAILED|Error|assert|mismatch|passed|failed" | head -30
