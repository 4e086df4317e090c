"""Shared helpers for the test-suite: golden loading, the host oracle library, comparisons."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
ORACLE_SO = os.path.join(ROOT, "oracle", "_build", "libnps_oracle.so")

# Tolerances from BASELINE.json: continuous state within 1e-9 relative per step and 1e-6 after 3 600 steps;
# discrete fields (flags, enums, counters, latches) bit-exact.
TOL_STEP = 1e-9
TOL_LONG = 1e-6
ABS_FLOOR = 1e-12   # |x| below this is compared absolutely (quantities that are exactly zero in steady state)


def oracle_lib() -> ctypes.CDLL:
    if not os.path.exists(ORACLE_SO):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")], stdout=subprocess.DEVNULL)
    L = ctypes.CDLL(ORACLE_SO)
    L.nps_oracle_n_state.restype = ctypes.c_int
    L.nps_oracle_n_params.restype = ctypes.c_int
    return L


def ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


def load_golden(name):
    from nuclear_sim_b200 import field_names
    z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    g = {k: z[k] for k in z.files}
    assert tuple(str(s) for s in g["state_names"]) == field_names("PlantState"), \
        f"{name}: fixture was generated for a different state.h (rerun oracle/make_golden.py)"
    assert tuple(str(s) for s in g["param_names"]) == field_names("PlantParams")
    return g


def load_sized(name):
    """64-plant BASELINE-size fixtures (oracle/make_golden_sized.py): the per-step inputs are regenerated from the plant
    ids by the generator's own pure function (no reference needed), everything else comes from the live reference."""
    from oracle import make_golden_sized as MG
    g = load_golden(name)
    T = int(g["n_steps"])
    ins = [MG.plant_inputs(name, j, int(pid), T) for j, pid in enumerate(g["plant_ids"])]
    g["actions"] = np.stack([i[0] for i in ins], axis=1)                      # [T, 64]
    g["magnitudes"] = np.stack([i[1] for i in ins], axis=1)
    g["noise"] = np.stack([i[2] for i in ins], axis=1)                        # [T, 64, 5]
    g["setpoint"] = np.full((T, len(ins)), np.nan)
    return g


def oracle_run_sized(L, g, st, t0, t1):
    """oracle_run for the sized fixtures: up to 4 injected (field, value) pairs per plant-step."""
    st = np.ascontiguousarray(st, dtype=np.float64).copy()
    inj = g["inject"]
    steps = sorted(set(np.nonzero(~np.isnan(inj[t0:t1, :, 0, 0]))[0] + t0)) + [t1]
    t = t0
    for nxt in steps:
        if nxt > t:
            st = oracle_run(L, st, g["params"], g["actions"], g["magnitudes"], g["noise"], g["setpoint"], None, t, nxt)
            t = nxt
        if t < t1:
            for p in np.nonzero(~np.isnan(inj[t, :, 0, 0]))[0]:
                for f, v in inj[t, p]:
                    if not np.isnan(f):
                        st[p, int(f)] = v
    return st


def discrete_mask():
    """Fields that hold flags / enums / counters / latches: compared bit-exactly.  The list is explicit: the members
    annotated `// @discrete` in csrc/plant/state.h (nuclear_sim_b200._layout.discrete_field_names)."""
    from nuclear_sim_b200 import field_names
    from nuclear_sim_b200._layout import discrete_field_names
    d = set(discrete_field_names())
    return np.array([n in d for n in field_names("PlantState")])


def canonicalize(v):
    """Rotate the pH deviation ring so its oldest entry sits at index 0 (storage detail, not state)."""
    from nuclear_sim_b200 import field_index
    ix = field_index()
    v = np.array(v, dtype=np.float64, copy=True)
    flat = v.reshape(-1, v.shape[-1])
    b = ix["ph.dev_hist[0]"]
    for row in flat:
        h = int(row[ix["ph.dev_head"]])
        if h:
            row[b:b + 100] = np.roll(row[b:b + 100], -h)
            row[ix["ph.dev_head"]] = 0.0
    return flat.reshape(v.shape)


def field_scales():
    """Natural magnitude below which a field is compared absolutely.  Default ABS_FLOOR; fields computed by
    catastrophic cancellation get their physical scale: tsp_flow_maldistribution is std/mean of seven nearly equal
    restrictions (a 0..1 factor tested against 0.30, tsp_fouling_model.py:369-392), so a 1-ulp libm difference
    in the restrictions is an absolute 1e-14 error on a quantity that may itself be 1e-7."""
    from nuclear_sim_b200 import field_names
    names = field_names("PlantState")
    sc = np.full(len(names), ABS_FLOOR)
    for i, n in enumerate(names):
        if n.endswith("tsp_flow_maldistribution"):
            sc[i] = 1.0
        # restriction = 1 - A_eff/A_orig with A_eff ~ A_orig for a clean plate (tsp_fouling_model.py:302-340): the
        # fouling fraction and everything proportional to it carry an absolute ~1e-16 rounding error, whatever
        # their size; their physical scale is 0..1 (stage thresholds 0.4 / 0.7 / 0.85).
        elif n.endswith(("tsp_fouling_fraction", "tsp_heat_transfer_degradation", "tsp_cumulative_power_loss")):
            sc[i] = 1e-6
    return sc


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    floor = field_scales() if a.shape[-1] == len(field_scales()) else ABS_FLOOR
    den = np.maximum(np.abs(b), floor)
    e = np.abs(a - b) / den
    e[a == b] = 0.0
    e[np.isnan(a) & np.isnan(b)] = 0.0
    return e


def assert_states_close(got, ref, tol, what=""):
    from nuclear_sim_b200 import field_names
    got = canonicalize(got)
    ref = canonicalize(ref)
    names = field_names("PlantState")
    dm = discrete_mask()
    bad_d = np.argwhere((got != ref) & dm[None, :] if got.ndim == 2 else (got != ref) & dm)
    assert bad_d.size == 0, f"{what}: discrete field mismatch, e.g. " + ", ".join(
        f"{names[i[-1]]} got {got[tuple(i)]} ref {ref[tuple(i)]}" for i in bad_d[:5])
    e = rel_err(got, ref)
    worst = np.unravel_index(np.argmax(e), e.shape)
    assert e.max() <= tol, (f"{what}: max rel err {e.max():.3e} > {tol:g} at {names[worst[-1]]} "
                            f"(got {got[worst]!r}, ref {ref[worst]!r})")
    return float(e.max())


def oracle_run(L, state0, params, actions, mags, noise, setpoint, inject, t0, t1):
    """Advance array-of-structs states [P, NS] from step t0 to t1 with the host oracle, honouring injections."""
    st = np.ascontiguousarray(state0, dtype=np.float64).copy()
    P = st.shape[0]
    params = np.ascontiguousarray(params, dtype=np.float64)
    for t in range(t0, t1):
        if inject is not None:
            for p in range(P):
                if not np.isnan(inject[t, p, 0]):
                    st[p, int(inject[t, p, 0])] = inject[t, p, 1]
        a = np.ascontiguousarray(actions[t], dtype=np.int8)
        m = np.ascontiguousarray(mags[t], dtype=np.float64)
        z = np.ascontiguousarray(noise[t], dtype=np.float64)
        sp = np.ascontiguousarray(setpoint[t], dtype=np.float64)
        rc = L.nps_oracle_step_sp(ptr(st), ptr(params), ptr(a), ptr(m), ptr(z), ptr(sp), ctypes.c_int64(P), 1)
        assert rc == 0
    return st


class OracleSim:
    """CPU stand-in for BatchedNuclearPlantSimulator built on the host oracle (TEST INFRASTRUCTURE): same surface as
    the device simulator for the maintenance host logic — step, threshold check with cooldown stamps
    (a numpy restatement of nps_threshold_kernel / state_manager.py:1307-1369), effects, value read-back."""

    def __init__(self, states, params):
        from nuclear_sim_b200 import field_index
        self.L = oracle_lib()
        self.L.nps_oracle_apply_maintenance.restype = ctypes.c_int
        self.st = np.ascontiguousarray(np.atleast_2d(states), dtype=np.float64).copy()
        self.params = np.ascontiguousarray(params, dtype=np.float64)
        self._initial = self.st.copy()
        self.n_plants = self.st.shape[0]
        self.ix = field_index()
        self._thr = None
        self._mon = None
        self.step_index = 0
        self.dt = float(self.params[field_index("PlantParams")["dt"]])

    # in-launch monitoring surface of BatchedNuclearPlantSimulator (enable_monitor / step(K=...) / drain_step_events)
    def enable_monitor(self, *a, **k):
        self._mon = True
        self._step_events = []

    def current_time_minutes(self):
        return float(self.st[:, self.ix["sim.time_minutes"]].max())

    def check_thresholds_events(self):
        self.check_thresholds()
        now = self.current_time_minutes()
        for p, t in self._fired:
            self._step_events.append((p, t, self.step_index - 1, 0, float(self._value(self._rows[t][0])[p]), now))
        self._fired = []

    def drain_step_events(self):
        from nuclear_sim_b200.batched import EVENT_DTYPE
        ev = np.array(sorted(self._step_events, key=lambda e: (e[2], e[0], e[1])), dtype=EVENT_DTYPE) if self._step_events \
            else np.zeros(0, dtype=EVENT_DTYPE)
        self._step_events = []
        return ev

    def step(self, actions=None, magnitudes=None, noise=None, power_setpoint=None, K=None, skip_last_check=False):
        if K is not None:     # device-style call: [K, N] inputs, noise [K, 5, N]; thresholds checked after every substep
            pick = lambda x, k: None if x is None else np.asarray(x)[k]   # noqa: E731
            for k in range(K):
                z = None if noise is None else np.ascontiguousarray(np.asarray(noise)[k].T)
                self.step(pick(actions, k), pick(magnitudes, k), z, pick(power_setpoint, k))
                if self._mon and getattr(self, "_rows", None) and not (skip_last_check and k == K - 1):
                    self.check_thresholds()
                    now = self.current_time_minutes()
                    for p, t in self._fired:
                        self._step_events.append((p, t, self.step_index - 1, 0, float(self._value(self._rows[t][0])[p]), now))
                    self._fired = []
            return
        self.step_index += 1
        P = self.n_plants
        a = np.full(P, 8, dtype=np.int8) if actions is None else np.ascontiguousarray(actions, dtype=np.int8)
        m = np.ones(P) if magnitudes is None else np.ascontiguousarray(magnitudes, dtype=np.float64)
        z = np.tile(np.array([0.0, 0.0, 1.0, 1.0, 1.0]), (P, 1)) if noise is None else np.ascontiguousarray(noise, dtype=np.float64)
        sp = np.full(P, np.nan) if power_setpoint is None else np.ascontiguousarray(power_setpoint, dtype=np.float64)
        assert self.L.nps_oracle_step_sp(ptr(self.st), ptr(self.params), ptr(a), ptr(m), ptr(z), ptr(sp), ctypes.c_int64(P), 1) == 0

    def state_numpy(self):
        return self.st.copy()

    # single-plant conveniences (same surface as BatchedNuclearPlantSimulator)
    def step_plant(self, plant, action, magnitude, z):
        assert self.n_plants == 1 and plant == 0
        self.step(actions=[action], magnitudes=[magnitude], noise=np.asarray(z, dtype=np.float64)[None, :])
        obs, rew = self._observe()
        return obs[0], float(rew[0]), bool(self.st[0, self.ix["pri.scram_activated"]] != 0.0)

    def _observe(self):
        obs = np.zeros((self.n_plants, 22)); rew = np.zeros(self.n_plants)
        self.L.nps_oracle_observe(ptr(np.ascontiguousarray(self.st)), ptr(self.params), ctypes.c_int64(self.n_plants), ptr(obs), ptr(rew))
        return obs, rew

    def observe_plant(self, plant):
        return self._observe()[0][plant]

    def reward_plant(self, plant):
        return float(self._observe()[1][plant])

    def reset_plant(self, plant):
        self.st[plant] = self._initial[plant]

    def write_fields(self, plant, values):
        for f, v in values.items():
            self.st[int(plant), int(f)] = float(v)

    # thresholds
    def set_thresholds(self, rows):
        code = {"greater_than": 0, "less_than": 1, "greater_equal": 2, "less_equal": 3, "equals": 4, "not_equals": 5}
        self._rows = [(-1 if r[0] is None else (int(r[0]) if isinstance(r[0], (int, np.integer)) else self.ix[r[0]]),
                       code.get(r[1], 6), float(r[2]), float(r[3]) * 60.0) for r in rows]
        self._last = np.full((len(self._rows), self.n_plants), -np.inf)
        self._fired = []

    def _value(self, f):
        if f >= 0:
            return self.st[:, f]
        k = -f - 2
        kind, unit = k >> 2, k & 3
        assert kind == 0
        w = lambda c: self.st[:, self.ix[f"fw.pump[{unit}].lub.component_wear[{c}]"]]
        return w(0) + np.maximum(np.maximum(w(1), w(2)), w(3)) + w(4)

    def check_thresholds(self):
        now = self.st[:, self.ix["sim.time_minutes"]]
        self._fired = []
        for t, (f, c, val, cd) in enumerate(self._rows):
            if f == -1:
                continue
            v = self._value(f)
            ready = ~((now - self._last[t]) < cd)
            with np.errstate(invalid="ignore"):
                fire = [v > val, v < val, v >= val, v <= val, np.abs(v - val) < 1e-3, np.abs(v - val) >= 1e-3,
                        np.zeros_like(v, dtype=bool)][c] & ready
            self._last[t, fire] = now[fire]
            self._fired += [(int(p), t) for p in np.nonzero(fire)[0]]

    def drain_events(self):
        return sorted(self._fired)

    def reset_cooldowns(self, plant, rows):
        self._last[list(rows), int(plant)] = -np.inf

    def read_threshold_values(self, plants, table):
        out = {}
        for t, (f, *_rest) in enumerate(self._rows):
            if f == -1:
                continue
            v = self._value(f)
            for p in plants:
                out[(p, t)] = float(v[p])
        return out

    def apply_maintenance(self, requests):
        status = []
        for plant, target, action, arg in requests:
            row = np.ascontiguousarray(self.st[plant]).copy()
            status.append(self.L.nps_oracle_apply_maintenance(ptr(row), ptr(self.params), int(target), int(action), int(arg)))
            self.st[plant] = row
        return status


def replay_maintenance_scenario(sim, golden, make_maintenance, check_state):
    """Drive `sim` through a maint_<scenario>.npz fixture with K=1 steps in the reference's order
    (physics -> maintenance update -> threshold check, sim.py:155-223); returns the BatchedAutoMaintenance."""
    import json
    log = json.loads(str(golden["log"]))
    from nuclear_sim_b200 import field_index
    dt = float(golden["params"][field_index("PlantParams")["dt"]])
    maint = make_maintenance(sim, log["maintenance_system"])
    T = golden["states"].shape[0]
    for t in range(T):
        z = golden["noise"][t][None, :] if golden["noise"].ndim == 2 else golden["noise"][t]     # [P, 5]
        if hasattr(sim, "slab"):      # device simulator: [K=1, 5, N] tensor
            import torch
            sim.step(noise=torch.from_numpy(np.ascontiguousarray(z.T[None])), K=1)
        else:
            sim.step(noise=z)
        t_min = (t + 1) * dt
        maint.update(t_min)
        maint.check(t_min)
        check_state(t, sim.state_numpy())
    return maint, log
