"""CPU: the device power function's ALGORITHMS (csrc/plant/fastpow.h: the series form the step kernel uses and the table-driven form
kept for tuning builds), compiled for the host by the oracle build purely for this test, against libm pow (numpy; glibc,
< 1 ulp) and against exact arithmetic (mpmath).  The plant step on the host never uses them; on the device the series
form replaces libdevice pow (<= 2 ulp) for positive finite bases."""
import ctypes
import math

import numpy as np

from tests import _util as U


import pytest

FORMS = ["nps_oracle_fastpow", "nps_oracle_fastpow_tab"]     # series (shipped), table-driven (tuning builds)
_FORM = ["nps_oracle_fastpow_tab"]


@pytest.fixture(params=FORMS, autouse=True)
def _form(request):
    _FORM[0] = request.param
    yield


def _fastpow(lib, x, y):
    x = np.ascontiguousarray(x, dtype=np.float64); y = np.ascontiguousarray(y, dtype=np.float64)
    out = np.zeros_like(x); taken = np.zeros(len(x), dtype=np.uint8)
    assert getattr(lib, _FORM[0])(U.ptr(x), U.ptr(y), U.ptr(out), taken.ctypes.data_as(ctypes.c_void_p), ctypes.c_int64(len(x))) == 0
    return out, taken.astype(bool)


def test_within_one_ulp_of_libm_on_model_operands(oracle_lib):
    rng = np.random.RandomState(1)
    n = 2_000_000
    x = 10.0 ** rng.uniform(-6, 6, n)
    y = rng.uniform(-3, 4, n)
    x[::7] = 1.0 + rng.uniform(-1e-3, 1e-3, len(x[::7]))
    y[::11] = rng.choice([0.38, 0.8, 0.15, -0.6, 1.8, 2.2, 2.4, 1.6, 1.4, 1.3, 1.0 / 3], len(y[::11]))
    got, taken = _fastpow(oracle_lib, x, y)
    assert taken.all()
    want = np.power(x, y)
    ulps = np.abs(got.view(np.int64) - want.view(np.int64))
    assert ulps.max() <= 1
    assert (ulps == 1).mean() < 0.15


def test_error_against_exact_arithmetic(oracle_lib):
    import mpmath as mp
    mp.mp.prec = 200
    rng = np.random.RandomState(2)
    x = 10.0 ** rng.uniform(-6, 6, 4000); y = rng.uniform(-3, 4, 4000)
    x[::5] = 1.0 + rng.uniform(-1e-2, 1e-2, len(x[::5]))          # near 1: e ln2 + log c cancels, the low words carry the result
    got, taken = _fastpow(oracle_lib, x, y)
    assert taken.all()
    worst = 0.0
    for xi, yi, gi in zip(x, y, got):
        t = mp.power(mp.mpf(float(xi)), mp.mpf(float(yi)))
        worst = max(worst, float(abs(mp.mpf(float(gi)) - t) / math.ulp(float(t))))
    assert worst < (0.52 if _FORM[0].endswith("_tab") else 1.2), worst


def test_guarded_range_refuses_what_it_cannot_do(oracle_lib):
    x = np.array([0.0, -0.0, -2.0, np.inf, np.nan, 5e-324, 1e-310, 2.0, 2.0, 2.0, 1e300, 1e-300, 1e5])
    y = np.array([1.8, 1.8, 3.0, 0.8, 1.2, 0.7, 0.7, np.inf, np.nan, 2000.0, 1.5, 1.5, 20.0])
    _, taken = _fastpow(oracle_lib, x, y)
    assert not taken.any()
    got, taken = _fastpow(oracle_lib, np.array([2.0, 1.0, 0.5, 10.0]), np.array([0.0, 3.3, 2.0, -2.0]))
    assert taken.all() and got.tolist() == [1.0, 1.0, 0.25, 0.01]
