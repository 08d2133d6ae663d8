import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    """Fixtures produced by tests/golden/make_golden.py from the reference's own source."""
    path = os.path.join(ROOT, "tests", "golden", "reference_vectors.npz")
    return dict(np.load(path))


def _f64(x):
    x = np.asarray(x)
    return x.astype(np.complex128) if np.iscomplexobj(x) else x.astype(np.float64)


def rel_fro(a, b):
    a, b = _f64(a), _f64(b)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def max_rel(a, b):
    a, b = _f64(a), _f64(b)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


@pytest.fixture(scope="session")
def golden_side():
    """Fixtures of the widened rows, produced by tests/golden/make_golden_side.py from the reference's own source."""
    return dict(np.load(os.path.join(ROOT, "tests", "golden", "reference_vectors_side.npz")))


@pytest.fixture(scope="session")
def golden_real():
    """Fixtures on the recordings the reference ships (img/gt_hfg.wav, img/y_tmpl.wav), produced by
    tests/golden/make_golden_real.py from the reference's own source.  Spectrogram outputs hold every 16th frame."""
    g = dict(np.load(os.path.join(ROOT, "tests", "golden", "reference_vectors_real.npz")))
    g["y_gt_hfg"] = g["wav_gt_hfg_int16"].astype(np.float32) / 32768.0
    g["y_y_tmpl"] = g["wav_y_tmpl"]
    for k in ("gt_hfg", "y_tmpl"):
        y = g[f"y_{k}"]
        g[f"y_{k}"] = y[:(len(y) // 256) * 256 - 1]           # aligned to the hop, then y[:-1] (retunegan/data.py:60-62)
    return g
