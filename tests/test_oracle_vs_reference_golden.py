"""CPU: pin the host restatement (oracle/cpu_port.cpp) against fixtures produced by stepping the live
Python reference (oracle/make_golden.py).  This is what makes the oracle trustworthy as the checker
for the CUDA path at sizes the reference cannot reach."""
import ctypes

import numpy as np
import pytest

from tests import _util as U

# trip_*: one plant each, a protection limit moved so that the normal operating point trips it (oracle/make_golden.py cfg8)
TRIP_SCENARIOS = ["trip_overspeed", "trip_vibration", "trip_bearing_temp", "trip_thrust_bearing", "trip_low_vacuum",
                  "trip_thermal_stress", "trip_rotor_vibration_alarms", "trip_vacuum_alarms", "trip_fw_low_suction",
                  "trip_fw_high_discharge", "trip_fw_low_flow", "trip_fw_npsh_alarm", "trip_fw_npsh_trip",
                  "trip_vacuum_lag_rotation", "trip_cond_tube_vibration", "trip_sg_no_load_balancing",
                  "trip_fw_manual_flow", "trip_rotor_slow", "trip_ejector_out_of_range"]
SCENARIOS = ["cfg1_oil_top_off", "cfg2_steady", "cfg3_loadfollow", "cfg4_scram", "cfg5_degradation", "cfg6_secondary_trips",
             "cfg7_turbine_trips_fouling", "cfg9_pump_trips_modes", "cfg10_primary_only", "cfg10_primary_only_constant"] + TRIP_SCENARIOS


@pytest.mark.parametrize("name", SCENARIOS)
def test_oracle_matches_reference_trajectory(oracle_lib, name):
    g = U.load_golden(name)
    st = g["state0"].copy()
    t = 0
    for c, cp in enumerate(g["checkpoints"]):
        st = U.oracle_run(oracle_lib, st, g["params"], g["actions"], g["magnitudes"], g["noise"], g["setpoint"],
                          g["inject"], t, int(cp))
        t = int(cp)
        tol = U.TOL_STEP * max(1, min(t, 1000)) if t < 3600 else U.TOL_LONG
        U.assert_states_close(st, g["states"][c], tol, f"{name} step {t}")


@pytest.mark.parametrize("name", SCENARIOS)
def test_oracle_observation_and_reward(oracle_lib, name):
    g = U.load_golden(name)
    P = g["state0"].shape[0]
    for c in range(len(g["checkpoints"])):
        st = np.ascontiguousarray(g["states"][c])
        obs = np.zeros((P, 22))
        rew = np.zeros(P)
        assert oracle_lib.nps_oracle_observe(U.ptr(st), U.ptr(np.ascontiguousarray(g["params"])), ctypes.c_int64(P),
                                             U.ptr(obs), U.ptr(rew)) == 0
        assert U.rel_err(obs, g["obs"][c]).max() <= 1e-12
        assert U.rel_err(rew, g["reward"][c]).max() <= 1e-9


def test_scram_steps_bit_exact(oracle_lib):
    g = U.load_golden("cfg4_scram")
    from nuclear_sim_b200 import field_index
    ix = field_index()
    st = g["state0"].copy()
    P = st.shape[0]
    first = np.full(P, -1)
    T = g["actions"].shape[0]
    for t in range(T):
        st = U.oracle_run(oracle_lib, st, g["params"], g["actions"], g["magnitudes"], g["noise"], g["setpoint"],
                          g["inject"], t, t + 1)
        act = st[:, ix["pri.scram_activated"]] != 0
        first[(first < 0) & act] = t
        np.testing.assert_allclose(st[:, ix["pri.power_level"]], g["power_level"][t], rtol=1e-9, atol=1e-12)
    assert first.tolist() == g["done_step"].tolist()
    assert (first >= 0).sum() >= 2   # the fixture really contains scrams


# ---- 64-plant fixtures built by the reference's own generators (BASELINE configs #3 and #4) ----------------------
def test_oracle_matches_reference_randomized_ics_cfg3(oracle_lib):
    """64 plants with the reference's get_randomized_*_conditions ICs over the action catalog, load-following / ramp /
    valve / boron policies, 3 600 steps: state at every checkpoint, power and electrical output every 10th step."""
    from nuclear_sim_b200 import field_index
    g = U.load_sized("cfg3_rand64")
    ix = field_index()
    tr = [ix[str(f)] for f in g["trace_fields"]]
    stride = int(g["trace_stride"])
    st = g["state0"].copy()
    t = 0
    cps = {int(c): i for i, c in enumerate(g["checkpoints"])}
    for row in range(g["trace"].shape[0]):
        t1 = (row + 1) * stride
        for c in sorted(k for k in cps if t < k < t1):
            st = U.oracle_run_sized(oracle_lib, g, st, t, c)
            t = c
            U.assert_states_close(st, g["states"][cps[c]], U.TOL_STEP * max(1, min(c, 1000)), f"cfg3_rand64 step {c}")
        st = U.oracle_run_sized(oracle_lib, g, st, t, t1)
        t = t1
        got = st[:, tr]
        assert np.array_equal(got[:, 0], g["trace"][row][:, 0])                 # scram_status
        assert U.rel_err(got, g["trace"][row]).max() <= (U.TOL_STEP * min(t, 1000) if t < 3600 else U.TOL_LONG), f"step {t}"
        if t in cps:
            tol = U.TOL_STEP * max(1, min(t, 1000)) if t < 3600 else U.TOL_LONG
            U.assert_states_close(st, g["states"][cps[t]], tol, f"cfg3_rand64 step {t}")
    assert t == 3600 and g["ic_randomized"].all()


def test_oracle_matches_reference_failure_injections_cfg4(oracle_lib):
    """64 plants, each with one EquipmentFailureSimulator / emergency transient at its own trigger step: scram_status,
    done, power, fuel temperature, flow, pressure EVERY step (event steps bit-exact), full state at the checkpoints."""
    from nuclear_sim_b200 import field_index
    g = U.load_sized("cfg4_fail64")
    ix = field_index()
    tr = [ix[str(f)] for f in g["trace_fields"]]
    st = g["state0"].copy()
    P = st.shape[0]
    first_scram = np.full(P, -1)
    first_done = np.full(P, -1)
    cps = {int(c): i for i, c in enumerate(g["checkpoints"])}
    for t in range(int(g["n_steps"])):
        st = U.oracle_run_sized(oracle_lib, g, st, t, t + 1)
        got = st[:, tr]
        assert np.array_equal(got[:, 0], g["trace"][t][:, 0]), f"scram_status at step {t}"
        done = st[:, ix["pri.scram_activated"]] != 0
        assert np.array_equal(done, g["done"][t].astype(bool)), f"done at step {t}"
        first_scram[(first_scram < 0) & (got[:, 0] != 0)] = t
        first_done[(first_done < 0) & done] = t
        assert U.rel_err(got, g["trace"][t]).max() <= U.TOL_STEP * (t + 1), f"step {t}"
        if t + 1 in cps:
            U.assert_states_close(st, g["states"][cps[t + 1]], U.TOL_STEP * (t + 1), f"cfg4_fail64 step {t + 1}")
    assert first_scram.tolist() == g["first_scram_step"].tolist()
    assert first_done.tolist() == g["first_done_step"].tolist()
    assert (first_scram >= 0).sum() >= 15
