"""Accuracy of the table-driven FP64 exp / log used in the likelihood epilogues."""
import ctypes

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _run(x, op):
    from causarray_b200 import _lib as L
    lib = L.load()
    x = np.ascontiguousarray(x, dtype=np.float64)
    out = np.empty_like(x)
    L.check(lib.ca_test_fastmath(L.dptr(x), x.size, op, L.dptr(out)))
    return out


def test_fm_exp():
    rng = np.random.default_rng(0)
    x = np.r_[rng.uniform(-700, 100, 200000), rng.uniform(-5, 5, 200000), [-700.0, 100.0, 0.0, -0.0, 1e-300, -1e-9]]
    got = _run(x, 0)
    ref = np.exp(x)
    rel = np.abs(got / ref - 1)
    print("fm_exp max rel err (ulps)", rel.max() / 2.220446049250313e-16)
    assert rel.max() <= 4 * 2.220446049250313e-16


def test_fm_log():
    rng = np.random.default_rng(1)
    x = np.r_[1 + rng.uniform(0, 1, 200000) ** 4, np.exp(rng.uniform(0, 100, 200000)), 1 + np.exp(rng.uniform(-40, 0, 50000)),
              [1.0, 2.0, 1e6, 1 + 1e-6]]
    got = _run(x, 1)
    ref = np.log(x)
    err = np.abs(got - ref)
    tol = 4 * 2.220446049250313e-16 * np.maximum(np.abs(ref), 2.0 ** -8)
    print("fm_log max err / tol", (err / tol).max())
    assert np.all(err <= tol)
