"""the drop-in TimingOptimizer class (nuclear_sim_b200.optimize.TimingOptimizer) against the
reference's own class (data_gen/optimization/timing_optimizer.py), loaded from /root/reference and run unmodified.

The reference's probe (_test_trigger_timing) builds one Python simulator per call; here BOTH classes are given the
same probe — the trigger time of the plant with the candidate initial condition, computed by the batched engine (host
oracle stand-in on CPU) — so the comparison is about the optimiser itself: which values it tries, how it moves its
bounds, what it returns.  (At HEAD the reference's own probe never sees a trigger: StateManager.record_maintenance_result
raises before the maintenance history is written)"""
import copy
import importlib.util
import json
import os

import numpy as np
import pytest

from tests import _util as U

try:
    from oracle import refplant as R
    HAVE_REF = R.reference_available()
except Exception:   # pragma: no cover
    HAVE_REF = False


def _fixture():
    g = np.load(os.path.join(U.GOLDEN, "maint_oil_top_off.npz"), allow_pickle=False)
    return g, json.loads(str(g["log"]))["maintenance_system"]


def _mini_config(mcfg):
    """A config with a handful of IC keys of the feedwater system (the optimiser walks every key it finds)."""
    return {"secondary_system": {"feedwater": {"initial_conditions": {
        "pump_oil_levels": [60.3, 60.3, 60.3, 60.3], "pump_oil_contamination": 8.0, "oil_temperature": 45.0,
        "control_mode": "auto", "pump_vibrations": [5.0, 1.0, 1.0, 0.0]}}}, "maintenance_system": mcfg}


def _make_probe(g, calls):
    from nuclear_sim_b200 import optimize as O
    fac = lambda st, p, dev: U.OracleSim(st, p)   # noqa: E731
    to = O.TimingOptimizer(verbose=False, base_state=g["state0"], params=g["params"], component_id=None, engine_factory=fac)

    def probe(configs, action, max_hours):
        calls.append(len(configs))
        to._probe = None
        try:
            return to._test_trigger_timing_batch(configs, action, max_hours)
        finally:
            to._probe = probe
    return probe


def test_batched_probe_reproduces_the_reference_trigger_time():
    """The probe: base plant of the maint_oil_top_off fixture -> first oil_top_off work order at 140 min, as the live
    reference logged it; lowering the initial oil level (IC key pump_oil_levels -> 4 state fields) triggers earlier."""
    from nuclear_sim_b200 import optimize as O
    g, mcfg = _fixture()
    log = json.loads(str(g["log"]))
    to = O.TimingOptimizer(verbose=False, base_state=g["state0"], params=g["params"],
                           engine_factory=lambda st, p, dev: U.OracleSim(st, p))
    cfg = _mini_config(mcfg)
    ic = cfg["secondary_system"]["feedwater"]["initial_conditions"]
    from nuclear_sim_b200 import field_index
    ic["pump_oil_levels"] = [float(g["state0"][field_index()["fw.pump[0].lub.oil_level"]])] * 4
    lower = copy.deepcopy(cfg)
    to._set_config_value(lower, "secondary_system.feedwater.initial_conditions.pump_oil_levels", 59.0)
    t = to._test_trigger_timing_batch([cfg, lower], "oil_top_off", 4.0)
    assert t[0] == log["created"][0]["t"] / 60.0
    assert t[1] is not None and t[1] < t[0]
    m = O.load_ic_field_map()
    assert m["secondary_system.feedwater.initial_conditions.pump_oil_levels"]["kind"] == "copy"
    assert sorted(m["secondary_system.feedwater.initial_conditions.pump_oil_levels"]["fields"]) == \
        sorted(f"fw.pump[{k}].lub.oil_level" for k in range(4))


@pytest.mark.skipif(not HAVE_REF, reason="live reference not present")
@pytest.mark.parametrize("target_hours,tol", [(1.5, 0.1), (0.75, 0.09)])
def test_same_result_as_the_reference_class(target_hours, tol):
    from nuclear_sim_b200 import optimize as O
    R.setup_paths()
    spec = importlib.util.spec_from_file_location(
        "ref_timing_optimizer", os.path.join(R.REF_ROOT, "nuclear_simulator", "data_gen", "optimization", "timing_optimizer.py"))
    mod = importlib.util.module_from_spec(spec)
    with R.quiet():
        spec.loader.exec_module(mod)
    g, mcfg = _fixture()
    cfg = _mini_config(mcfg)
    ref_calls, our_calls = [], []
    probe_ref = _make_probe(g, ref_calls)
    ref = mod.TimingOptimizer(verbose=False)
    ref._test_trigger_timing = lambda config, action, max_hours: probe_ref([config], action, max_hours)[0]
    want = ref.optimize_for_target_timing(copy.deepcopy(cfg), "oil_top_off", target_hours, tolerance_hours=tol, max_iterations=10)
    ours = O.TimingOptimizer(verbose=False, probe=_make_probe(g, our_calls))
    got = ours.optimize_for_target_timing(copy.deepcopy(cfg), "oil_top_off", target_hours, tolerance_hours=tol, max_iterations=10)
    assert got[0] == want[0]                      # the optimised config, every key
    assert got[1] == want[1] and got[2] == want[2]
    assert want[1] is not None and abs(want[1] - target_hours) <= max(tol, 1.0 / 60.0) + 1e-12
    # helper parity on every IC key of the composed template config
    full = R.compose_config("oil_top_off")
    assert ours._extract_initial_conditions(full, "x") == ref._extract_initial_conditions(full, "x")
    for path, v in ref._extract_initial_conditions(full, "x").items():
        assert ours._get_parameter_bounds(path, v) == ref._get_parameter_bounds(path, v)
        assert ours._parameter_increases_degradation(path) == ref._parameter_increases_degradation(path)
    # the batched form needs one engine run per parameter searched, the sequential form one per probe
    assert len(our_calls) < len(ref_calls) and max(our_calls) > 1
