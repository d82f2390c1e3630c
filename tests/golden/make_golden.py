"""Generates tests/golden/square_test_expected.npy.

The reference's own CPU path for its one known-answer test is `get_non_dirt_pixels()` in
/root/reference/tests/square_test.py:11-17 (pure TensorFlow meshgrid arithmetic; TensorFlow is not installed
in this image, so the seven lines are restated with numpy in oracle/numpy_oracle.square_reference_pixels).
Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', '..'))
from oracle import numpy_oracle  # noqa: E402

if __name__ == '__main__':
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'square_test_expected.npy')
    np.save(out, numpy_oracle.square_reference_pixels(128, 128, 32, 64, 16))
    print('wrote', out)
