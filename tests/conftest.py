"""pytest configuration: `gpu` marker, seeded rng (same seed as the reference's tests/conftest.py:21-24),
golden-fixture loader."""

from __future__ import annotations

import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture
def rng():
    return np.random.default_rng(seed=42)


class Golden:
    """Access `case/attr` arrays of one tests/golden/<group>_<precision>.npz file."""

    def __init__(self, group: str, precision: str):
        self._z = np.load(os.path.join(GOLDEN_DIR, f"{group}_{precision}.npz"))

    def case(self, name: str) -> dict:
        pre = name + "/"
        out = {k[len(pre):]: self._z[k] for k in self._z.files if k.startswith(pre)}
        if not out:
            raise KeyError(name)
        return out


def load_golden(group: str, precision: str) -> Golden:
    return Golden(group, precision)


def real_t_of(precision: str):
    return np.float32 if precision == "single" else np.float64


def test_tol(precision: str) -> float:
    """1e3 * eps, the reference's get_test_tol (sopht/utils/precision.py:16-19)."""
    t = real_t_of(precision)
    return float(t(1e3) * np.finfo(t).eps)


def rel_l2(a, b) -> float:
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    den = np.linalg.norm(b.ravel())
    return float(np.linalg.norm((a - b).ravel()) / (den if den > 0 else 1.0))


# BASELINE.json north_star: per-kernel relative L2 error <= 1e-5 (fp32) / 1e-12 (fp64)
REL_L2_TOL = {"single": 1e-5, "double": 1e-12}
