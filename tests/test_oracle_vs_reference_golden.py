"""CPU: pin the host restatement (oracle/cpu_port.cpp) against fixtures produced by stepping the live
Python reference (oracle/make_golden.py).  This is what makes the oracle trustworthy as the checker
for the CUDA path at sizes the reference cannot reach."""
import ctypes

import numpy as np
import pytest

from tests import _util as U

SCENARIOS = ["cfg1_oil_top_off", "cfg2_steady", "cfg3_loadfollow", "cfg4_scram", "cfg5_degradation", "cfg6_secondary_trips", "cfg7_turbine_trips_fouling"]


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
