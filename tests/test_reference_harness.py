"""CPU, needs /root/reference (skipped on the GPU box): the oracle harness itself — the host-supplied
random streams reproduce the reference's own RNG calls bit for bit, and one live reference step equals
the host restatement."""
import ctypes

import numpy as np
import pytest

from tests import _util as U

try:
    from oracle import refplant as R
    HAVE_REF = R.reference_available()
except Exception:   # pragma: no cover
    HAVE_REF = False

pytestmark = pytest.mark.skipif(not HAVE_REF, reason="live reference not present")


def test_stream_normal_equals_randomstate_normal():
    """RandomState(seed).normal(0, s) == 0 + s * RandomState(seed).standard_normal()
    (constant_heat_source.py:60,178; tests/test_noise_heat_source.py:157-233 pins same-seed reproducibility)."""
    a, b = np.random.RandomState(42), np.random.RandomState(42)
    for s in (0.3, 2.7, 3000 * 0.001):
        for _ in range(200):
            assert a.normal(0.0, s) == 0.0 + s * b.standard_normal()


def test_reactivity_known_answers():
    """Range pins from the reference's own unit tests (tests/test_reactivity_model.py:98,189,209,360-366)."""
    R.setup_paths()
    with R.quiet():
        from systems.primary.reactor.reactivity_model import ReactivityModel, create_equilibrium_state
    m = ReactivityModel()
    assert -15000 <= m.calculate_boron_reactivity(1200.0) <= -10000
    eq = create_equilibrium_state()
    total, comp = m.calculate_total_reactivity(eq)
    # the reference's own test pins |total| <= 50 pcm (tests/test_reactivity_model.py:360-366); at HEAD the
    # equilibrium state sits at -100 pcm (rods at 95 % => +1350, boron balanced with the rods then moved) — the
    # reference test is stale, so the value the live code actually produces is what is pinned here.
    assert abs(total - (-100.0)) < 1e-6
    assert -3000 <= comp["xenon"] <= -1000 and -1000 <= comp["samarium"] <= -300


def test_live_reference_step_equals_restatement(oracle_lib):
    rp = R.make_reference_plant(R.compose_config("oil_top_off"), dt=5.0, heat_source="constant", noise_enabled=True,
                                noise_std_percent=0.1)
    params = R.extract_params(rp.sim)
    rng = np.random.RandomState(3)
    for k in range(5):
        s0 = R.extract_state(rp.sim)
        z = np.array([rng.standard_normal(), rng.standard_normal(), rng.random_sample(), rng.random_sample(), rng.random_sample()])
        rp.step(8, 1.0, z)
        s1 = R.extract_state(rp.sim)
        c = s0.copy()[None, :]
        assert oracle_lib.nps_oracle_step(U.ptr(c), U.ptr(params), U.ptr(np.array([8], dtype=np.int8)), U.ptr(np.array([1.0])),
                                          U.ptr(z), ctypes.c_int64(1), 1) == 0
        U.assert_states_close(c, s1[None, :], 1e-12, f"live step {k}")


def test_long_format_export_equals_reference_plant_data_logger(tmp_path):
    """the long-format trajectory CSV.  The reference's own PlantDataLogger
    (data/plant_data_logger.py:82-157) logs a live plant for 12 steps; export_long_format writes the same steps from the
    extracted PlantState vectors.  Every row — parameter name, value text, unit, quality, and their order — must be equal
    (the timestamp column is datetime.now() in the reference, so it is copied over)."""
    import csv
    import os
    import sys
    from nuclear_sim_b200 import export as E
    rp = R.make_reference_plant(R.compose_config("oil_top_off"), dt=1.0, heat_source="reactor")
    sys.path.insert(0, os.path.join(R.REF_ROOT, "data"))
    with R.quiet():
        from plant_data_logger import PlantDataLogger
    logger = PlantDataLogger(str(tmp_path / "run"))
    states, times = [], []
    rng = np.random.RandomState(3)
    for t in range(12):
        rp.step(int(rng.choice([0, 1, 8, 9, 10])), 0.7, (0.0, 0.0, 1.0, 1.0, 1.0))
        logger.log_timestep(rp.sim)
        states.append(R.extract_state(rp.sim))
        times.append(float(rp.sim.time))
    ref_rows = list(csv.reader(open(logger.csv_path, newline="")))
    stamps = [ref_rows[1 + 22 * i][0] for i in range(12)]
    out = tmp_path / "ours.csv"
    n = E.export_long_format(str(out), stamps, states, times)
    got_rows = list(csv.reader(open(out, newline="")))
    assert n == 12 * 22 and len(got_rows) == len(ref_rows) == 1 + 12 * 22
    assert got_rows == ref_rows
