import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with `-m gpu`)')


@pytest.fixture(scope='session')
def oracle():
    from oracle import oracle as o
    o.build()
    return o


@pytest.fixture(scope='session')
def cuda_lib():
    """Builds (if stale) and loads libdirt_b200.so; GPU tests call the product through it."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    from dirt_b200 import build, _lib
    build.build()
    return _lib.lib()


WORST_RATIOS = {}   # name -> largest (abs err / allowed) any test saw: the margin left under the tolerance


def pytest_terminal_summary(terminalreporter):
    if WORST_RATIOS:
        terminalreporter.write_line('worst error / tolerance ratio per tensor (1.0 = at the bar of 1e-4 relative):')
        for name in sorted(WORST_RATIOS):
            terminalreporter.write_line('  %-28s %.3f' % (name, WORST_RATIOS[name]))


def rel_close(actual, expected, rel=1e-4, scale_frac=1e-2, name=None):
    """abs(a-b) <= rel * max(|a|, |b|, scale) with scale = scale_frac * max|expected|.
    Returns (ok, worst_ratio) where ratio = abs err / allowed; with `name`, the ratio also goes into the end-of-run
    summary (normalised to rel = 1e-4)."""
    a = np.asarray(actual, np.float64)
    b = np.asarray(expected, np.float64)
    scale = scale_frac * (np.abs(b).max() if b.size else 0.0)
    allowed = rel * np.maximum(np.maximum(np.abs(a), np.abs(b)), max(scale, 1e-30))
    err = np.abs(a - b)
    ratio = float((err / allowed).max()) if a.size else 0.0
    if name is not None:
        WORST_RATIOS[name] = max(WORST_RATIOS.get(name, 0.0), ratio * rel / 1e-4)
    return ratio <= 1.0, ratio
