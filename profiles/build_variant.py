"""Builds a tuning variant of the library next to the product one:  python profiles/build_variant.py TAG [-DNAME=VALUE ...]
-> dirt_b200/variants/libdirt_b200_TAG.so (git-ignored like every .so; it travels to the GPU box with the snapshot)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dirt_b200 import build as b  # noqa: E402

tag, extra = sys.argv[1], sys.argv[2:]
out_dir = os.path.join(ROOT, 'dirt_b200', 'variants')
os.makedirs(out_dir, exist_ok=True)
out = os.path.join(out_dir, 'libdirt_b200_%s.so' % tag)
srcs = [os.path.join(ROOT, 'dirt_b200', 'csrc', s) for s in b.SOURCES]
cmd = [b._nvcc()] + b.NVCC_FLAGS + extra + ['-o', out] + srcs
proc = subprocess.run(cmd, capture_output=True, text=True)
if proc.returncode != 0:
    sys.exit(proc.stdout + proc.stderr)
print(out)
