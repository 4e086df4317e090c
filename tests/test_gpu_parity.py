"""GPU: the CUDA path, called through the C ABI, against (a) the committed reference fixtures and
(b) the host oracle on seeded batches; plus size-independent properties at BASELINE sizes."""
import ctypes

import os

import numpy as np
import pytest

from tests import _util as U

pytestmark = pytest.mark.gpu

from tests.test_oracle_vs_reference_golden import TRIP_SCENARIOS  # noqa: E402  (one plant per protection path)

SCENARIOS = ["cfg1_oil_top_off", "cfg2_steady", "cfg3_loadfollow", "cfg4_scram", "cfg5_degradation", "cfg6_secondary_trips",
             "cfg7_turbine_trips_fouling", "cfg9_pump_trips_modes", "cfg10_primary_only", "cfg10_primary_only_constant"] + TRIP_SCENARIOS


def _sim(state0, params):
    from nuclear_sim_b200 import BatchedNuclearPlantSimulator
    return BatchedNuclearPlantSimulator(state0.shape[0], state0, params, device="cuda:0")


def _advance(sim, g, t0, t1, kmax=64):
    """Advance with fused substeps, splitting launches at injection points."""
    import torch
    t = t0
    inj = g["inject"]
    while t < t1:
        for p in range(sim.n_plants):
            if not np.isnan(inj[t, p, 0]):
                sim.slab[int(inj[t, p, 0]), p] = float(inj[t, p, 1])
        k = 1
        while t + k < t1 and k < kmax and np.isnan(inj[t + k, :, 0]).all():
            k += 1
        sim.step(actions=torch.from_numpy(np.ascontiguousarray(g["actions"][t:t + k])),
                 magnitudes=torch.from_numpy(np.ascontiguousarray(g["magnitudes"][t:t + k])),
                 noise=torch.from_numpy(np.ascontiguousarray(g["noise"][t:t + k].transpose(0, 2, 1))),
                 power_setpoint=torch.from_numpy(np.ascontiguousarray(g["setpoint"][t:t + k])), K=k)
        t += k


@pytest.mark.parametrize("name", SCENARIOS)
def test_cuda_matches_reference_fixture(name):
    g = U.load_golden(name)
    sim = _sim(g["state0"], g["params"])
    t = 0
    for c, cp in enumerate(g["checkpoints"]):
        _advance(sim, g, t, int(cp))
        t = int(cp)
        tol = U.TOL_STEP * max(1, min(t, 1000)) if t < 3600 else U.TOL_LONG
        U.assert_states_close(sim.state_numpy(), g["states"][c], tol, f"{name} step {t}")
        obs = sim.get_observation().cpu().numpy()
        assert U.rel_err(obs, g["obs"][c]).max() <= 1e-9


def test_cuda_scram_steps_bit_exact():
    import torch
    g = U.load_golden("cfg4_scram")
    sim = _sim(g["state0"], g["params"])
    T = g["actions"].shape[0]
    first = np.full(sim.n_plants, -1)
    for t in range(T):
        for p in range(sim.n_plants):
            if not np.isnan(g["inject"][t, p, 0]):
                sim.slab[int(g["inject"][t, p, 0]), p] = float(g["inject"][t, p, 1])
        out = sim.step(actions=torch.from_numpy(g["actions"][t]), magnitudes=torch.from_numpy(g["magnitudes"][t]),
                       noise=torch.from_numpy(np.ascontiguousarray(g["noise"][t].T)))
        d = out["done"].cpu().numpy()
        first[(first < 0) & d] = t
        np.testing.assert_allclose(sim.state.power_level.cpu().numpy(), g["power_level"][t], rtol=1e-9, atol=1e-12)
    assert first.tolist() == g["done_step"].tolist()


@pytest.mark.parametrize("n,steps,k", [(1, 20, 1), (33, 16, 4), (4096, 12, 6)])
def test_cuda_matches_host_oracle_on_random_batch(oracle_lib, n, steps, k):
    """Seeded, ragged batch sizes (1, non-multiple-of-warp, config-#2 size): perturbed ICs, random actions."""
    import torch
    from nuclear_sim_b200 import field_index, load_snapshot
    s0, params = load_snapshot("pwr3000_reactor_dt1")
    ix = field_index()
    rng = np.random.RandomState(123 + n)
    st = np.tile(s0, (n, 1))
    for f, lo, hi in (("pri.control_rod_position", 80, 100), ("pri.fuel_temperature", 540, 620),
                      ("fw.pump[0].lub.oil_level", 20, 95), ("fw.pump[1].lub.component_wear[2]", 0, 8),
                      ("sgs.sg[0].tif_scale_thickness", 0, 1.0), ("turb.lub.oil_contamination_level", 2, 12),
                      ("cond.fl_biofouling_thickness", 0, 1.0), ("sgs.sg[1].water_level", 11.5, 13.5)):
        st[:, ix[f]] = rng.uniform(lo, hi, n)
    acts = rng.choice([0, 1, 2, 3, 4, 5, 8, 9, 10], size=(steps, n)).astype(np.int8)
    mags = rng.uniform(0.2, 1.0, (steps, n))
    noise = np.stack([rng.standard_normal((steps, n)), rng.standard_normal((steps, n)), rng.random_sample((steps, n)),
                      rng.random_sample((steps, n)), rng.random_sample((steps, n))], axis=2)   # [T, n, 5]
    setp = np.full((steps, n), np.nan)
    ref = U.oracle_run(oracle_lib, st, params, acts, mags, noise, setp, None, 0, steps)
    sim = _sim(st, params)
    for t in range(0, steps, k):
        sim.step(actions=torch.from_numpy(acts[t:t + k]), magnitudes=torch.from_numpy(mags[t:t + k]),
                 noise=torch.from_numpy(np.ascontiguousarray(noise[t:t + k].transpose(0, 2, 1))), K=k)
    U.assert_states_close(sim.state_numpy(), ref, U.TOL_STEP * steps, f"n={n}")


def test_fused_substeps_equal_single_steps():
    """K fused substeps in one launch == K launches of one step (bitwise)."""
    import torch
    from nuclear_sim_b200 import load_snapshot
    s0, params = load_snapshot("pwr3000_oil_top_off_dt5")
    n, K = 257, 8
    rng = np.random.RandomState(5)
    noise = torch.from_numpy(np.stack([rng.standard_normal((K, n)), rng.standard_normal((K, n)), rng.random_sample((K, n)),
                                       rng.random_sample((K, n)), rng.random_sample((K, n))], axis=1))
    a = _sim(np.tile(s0, (n, 1)), params)
    b = _sim(np.tile(s0, (n, 1)), params)
    a.step(noise=noise, K=K)
    for k in range(K):
        b.step(noise=noise[k], K=1)
    assert torch.equal(a.slab, b.slab)


def test_identical_plants_stay_identical_at_65536():
    """BASELINE size property: plants are independent, so identical ICs + identical inputs give identical
    trajectories for every plant, and permuting plants permutes results."""
    import torch
    from nuclear_sim_b200 import load_snapshot
    s0, params = load_snapshot("pwr3000_reactor_dt1")
    n = 65536
    sim = _sim(np.tile(s0, (n, 1)), params)
    act = torch.full((4, n), 1, dtype=torch.int8)
    sim.step(actions=act, K=4)
    assert bool((sim.slab == sim.slab[:, :1]).all())
    one = _sim(s0[None, :], params)
    one.step(actions=torch.full((4, 1), 1, dtype=torch.int8), K=4)
    assert torch.equal(sim.slab[:, 12345], one.slab[:, 0])


def test_threshold_flags_and_ring_buffer():
    import torch
    from nuclear_sim_b200 import field_index, load_snapshot
    s0, params = load_snapshot("pwr3000_oil_top_off_dt5")
    ix = field_index()
    n = 1000
    st = np.tile(s0, (n, 1))
    rng = np.random.RandomState(0)
    st[:, ix["fw.pump[0].lub.oil_level"]] = rng.uniform(55, 65, n)
    sim = _sim(st, params)
    rows = [("fw.pump[0].lub.oil_level", "<", 58.0, 168.0), (None, ">", 1.0, 1.0),
            ("sgs.sg[0].tsp_fouling_fraction", ">", 0.3, 1.0), ("pri.power_level", ">=", 50.0, 0.25)]
    sim.set_thresholds(rows)
    sim.set_logged_fields(["pri.power_level", "fw.pump[0].lub.oil_level", "sec.electrical_power_output"], ring_rows=4)
    fired_total = np.zeros(n, dtype=int)
    for step in range(6):
        sim.step()
        sim.log_row()
        flags, anyw = sim.check_thresholds()
        ev = sim.drain_events()
        lvl = sim.state["fw.pump[0].lub.oil_level"].cpu().numpy()
        f0 = {p for p, t in ev if t == 0}
        fired_total[list(f0)] += 1
        # a plant fires threshold 0 exactly once (cooldown 168 h) and only when below 58 %
        assert all(lvl[p] < 58.0 for p in f0)
        assert not any(t == 1 for _, t in ev)            # unbound threshold is inert
        f3 = {p for p, t in ev if t == 3}
        # power >= 50 with a 15-minute cooldown at dt = 5 min fires on steps 0, 3
        assert (len(f3) == n) == (step % 3 == 0)
    lvl = sim.state["fw.pump[0].lub.oil_level"].cpu().numpy()
    assert ((fired_total == 1) == (lvl < 58.0)).all() and fired_total.max() == 1
    log = sim.drain_log()
    assert log.shape == (4, 3, n)
    np.testing.assert_array_equal(log[-1, 1], lvl)
    np.testing.assert_array_equal(log[-1, 0], sim.state.power_level.cpu().numpy())


@pytest.mark.parametrize("n", [2, 1000, 2050, 4098, 1001, 4097])
def test_ring_buffer_row_is_a_bitwise_gather(n):
    """nps_log_row on both of its kernels: even plant counts take the TMA bulk-copy pipeline (2050 / 4098: a ragged last
    16 KB chunk; 2: a 16-byte transfer), odd ones the shared-memory tile kernel.  Every logged field, every plant, every
    ring slot must be the slab's bits; the ring wraps after `rows` writes."""
    import torch
    from nuclear_sim_b200 import field_names, load_snapshot
    s0, params = load_snapshot("pwr3000_oil_top_off_dt5")
    rng = np.random.RandomState(n)
    st = np.tile(s0, (n, 1)) * (1.0 + 1e-3 * rng.standard_normal((n, 1)))
    sim = _sim(st, params)
    names = list(field_names())
    picked = [names[i] for i in sorted(rng.choice(len(names), size=97, replace=False))]
    sim.set_logged_fields(picked, ring_rows=3)
    want = []
    for step in range(5):
        sim.step()
        sim.log_row()
        want.append(np.stack([sim.state[f].cpu().numpy() for f in picked]))
    torch.cuda.synchronize()
    log = sim.drain_log()
    assert log.shape == (3, 97, n)
    for got, ref in zip(log, want[-3:]):
        np.testing.assert_array_equal(got.view(np.uint64), ref.view(np.uint64))


# ---- maintenance: flag kernel -> work orders -> effects on the device state --------------------------------------
MAINT_SCENARIOS = ["oil_top_off", "tsp_chemical_cleaning", "oil_change", "scale_removal"]


@pytest.mark.parametrize("name", MAINT_SCENARIOS)
def test_cuda_maintenance_scenario_matches_reference(name):
    """Full loop through the C ABI (nps_step, nps_check_thresholds, nps_apply_maintenance) against the live-reference
    fixture: identical threshold events / work orders / execution steps, state within tolerance at every step."""
    import os
    from nuclear_sim_b200 import maintenance as M
    from tests.test_maintenance_host import compare_logs
    g = np.load(os.path.join(U.GOLDEN, f"maint_{name}.npz"), allow_pickle=False)
    sim = _sim(g["state0"][None, :], g["params"])

    def check_state(t, got):
        U.assert_states_close(got, g["states"][t][None, :], U.TOL_STEP * (t + 1), f"{name} step {t}")
    maint, log = U.replay_maintenance_scenario(
        sim, g, lambda s, cfg: M.BatchedAutoMaintenance(s, M.ThresholdTable(cfg), aggressive=True), check_state)
    compare_logs(maint, log)


def test_cuda_maintenance_effects_match_reference():
    """nps_apply_maintenance for every (component, action) fixture pair, batched: one plant per case."""
    import os
    from nuclear_sim_b200 import maintenance as M
    z = np.load(os.path.join(U.GOLDEN, "maint_effects.npz"), allow_pickle=False)
    n = len(z["component"])
    sim = _sim(np.ascontiguousarray(z["before"]), z["params"])
    req = []
    for i in range(n):
        action, sub = str(z["action"][i]), str(z["arg"][i])
        req.append((i, M.target_code(str(z["component"][i])), M.action_code(action),
                    M.BEARING_ARG.get(sub, 0) if action == "bearing_replacement" else 0))
    status = sim.apply_maintenance(req)
    assert [bool(s == 1) for s in status] == [bool(b) for b in z["success"]]
    U.assert_states_close(sim.state_numpy(), z["after"], 1e-13, "maintenance effects")


def test_cuda_catalogue_sweep_matches_reference():
    """Every catalogued action name x every addressable component class (tests/golden/maint_effects_sweep.npz, 3 014 calls
    of the live reference): one plant per call, started from the state the reference had before that call, all applied
    in ONE nps_apply_maintenance; success flags and resulting states must equal the reference's."""
    from nuclear_sim_b200 import maintenance as M
    from tests.test_maintenance_host import _load_sweep, sweep_cases
    z = _load_sweep()
    cases = list(sweep_cases(z))
    before = np.ascontiguousarray(np.stack([c[3] for c in cases]))
    after = np.stack([c[4] for c in cases])
    sim = _sim(before, z["params"])
    status = sim.apply_maintenance([(i, M.target_code(c[0]), M.action_code(c[1]), 0) for i, c in enumerate(cases)])
    bad = [(c[0], c[1]) for c, s in zip(cases, status) if bool(s == 1) != c[2]]
    assert not bad, f"success flag differs from the reference for {bad[:5]}"
    U.assert_states_close(sim.state_numpy(), after, 1e-13, "catalogue sweep")


def test_cuda_maintenance_batched_equals_oracle_host_logic():
    """256 plants with staggered oil levels / fouling: the device loop and the oracle stand-in must issue the same
    work orders for the same plants at the same steps, and end in the same state."""
    import os
    import json
    from nuclear_sim_b200 import maintenance as M, field_index
    g = np.load(os.path.join(U.GOLDEN, "maint_oil_top_off.npz"), allow_pickle=False)
    cfg = json.loads(str(g["log"]))["maintenance_system"]
    ix = field_index()
    n = 256
    st = np.tile(g["state0"], (n, 1))
    rng = np.random.RandomState(5)
    st[:, ix["fw.pump[0].lub.oil_level"]] = 58.0 + rng.uniform(0.0, 1.5, n)
    st[:, ix["fw.pump[2].lub.oil_level"]] = 58.0 + rng.uniform(0.0, 3.0, n)
    st[:, ix["fw.pump[1].lub.oil_contamination_level"]] = 15.2 - rng.uniform(0.0, 0.02, n)
    dt = float(g["params"][field_index("PlantParams")["dt"]])
    sims = [_sim(st, g["params"]), U.OracleSim(st, g["params"])]
    maints = [M.BatchedAutoMaintenance(s, M.ThresholdTable(cfg), aggressive=True) for s in sims]
    import torch
    for t in range(30):
        sims[0].step(K=1)
        sims[1].step()
        for m in maints:
            m.update((t + 1) * dt)
            m.check((t + 1) * dt)
    key = lambda w: (w.plant, w.created, w.component_id, w.action, w.priority, w.executed_at, w.success)
    assert [key(w) for w in maints[0].created_log] == [key(w) for w in maints[1].created_log]
    assert len(maints[0].created_log) > 50 and len(maints[0].executed_log) > 50
    U.assert_states_close(sims[0].state_numpy(), sims[1].state_numpy(), U.TOL_STEP * 30, "batched maintenance")


@pytest.mark.parametrize("cls", ["ColumnarAutoMaintenance", "NativeAutoMaintenance"])
def test_cuda_columnar_maintenance_equals_object_bookkeeping(cls):
    """The large-batch path end to end on the device: in-launch threshold events, event-list flag kernel at the gate
    steps, array bookkeeping (numpy columns / the library's native work-order table), effects applied by the maintenance kernel — against the object bookkeeping on the
    host oracle stand-in, 512 plants staggered around several thresholds, 36 steps."""
    import json
    from nuclear_sim_b200 import maintenance as M, field_index
    g = np.load(os.path.join(U.GOLDEN, "maint_oil_top_off.npz"), allow_pickle=False)
    cfg = json.loads(str(g["log"]))["maintenance_system"]
    ix = field_index()
    n = 512
    st = np.tile(g["state0"], (n, 1))
    rng = np.random.RandomState(21)
    st[:, ix["fw.pump[0].lub.oil_level"]] = 58.0 + rng.uniform(-0.3, 1.5, n)
    st[:, ix["fw.pump[2].lub.oil_level"]] = 58.0 + rng.uniform(0.0, 3.0, n)
    st[:, ix["fw.pump[1].lub.oil_contamination_level"]] = 15.2 - rng.uniform(-0.01, 0.03, n)
    st[::3, ix["fw.pump[3].lub.component_wear[1]"]] = 8.6
    st[::5, ix["fw.pump[3].lub.component_wear[4]"]] = 16.5
    dev, host = _sim(st, g["params"]), U.OracleSim(st, g["params"])
    md = getattr(M, cls)(dev, M.ThresholdTable(cfg), aggressive=True)
    mh = M.BatchedAutoMaintenance(host, M.ThresholdTable(cfg), aggressive=True)
    md.advance(36)
    mh.advance(36)
    md.materialize_logs()
    key = lambda w: (w.created, w.plant, w.component_id, w.action, w.work_order_id, w.priority, w.sub_component, w.executed_at, w.success)   # noqa: E731
    assert sorted(key(w) for w in md.created_log) == sorted(key(w) for w in mh.created_log)
    assert md.n_work_orders_created > 600 and md.n_work_orders_executed > 400
    U.assert_states_close(dev.state_numpy(), host.state_numpy(), U.TOL_STEP * 36, "columnar maintenance on the device")


def test_cuda_scalar_facade_runs_maintenance_scenario(tmp_path):
    """The reference-API facade (plant_simulator.NuclearPlantSimulator) on a 1-plant CUDA engine: step() dicts, work
    orders through the runner-facing members, CSV export in the reference schema."""
    import csv
    import json
    import os
    import types
    from nuclear_sim_b200.plant_simulator import NuclearPlantSimulator
    g = np.load(os.path.join(U.GOLDEN, "maint_oil_top_off.npz"), allow_pickle=False)
    log = json.loads(str(g["log"]))
    sim = NuclearPlantSimulator(dt=5.0, initial_state=g["state0"], params=g["params"], device="cuda:0")
    sim.state_manager.config = {"maintenance_system": log["maintenance_system"]}
    sim.maintenance_system.setup_monitoring_from_state_manager(sim.state_manager, aggressive_mode=True)
    created = []
    for t in range(g["states"].shape[0]):
        z = g["noise"][t]
        sim._ph_rng = types.SimpleNamespace(standard_normal=lambda z=z: float(z[1]),
                                            random_sample=lambda z=z, it=iter([2, 3, 4]): float(z[next(it)]))
        out = sim.step(action=None)
        assert out["observation"].shape == (22,) and isinstance(out["reward"], float) and out["done"] is False
        if sim.maintenance_system.current_update_work_orders:
            created.append(t)
        U.assert_states_close(sim._engine.state_numpy(), g["states"][t][None, :], U.TOL_STEP * (t + 1), f"facade step {t}")
    assert created == [c["step"] for c in log["created"]]
    fwp = tmp_path / "fwp.csv"
    sim.state_manager.export_by_subcategory("secondary", "feedwater_FWP-1", str(fwp))
    rows = list(csv.reader(open(fwp)))
    j = rows[0].index("secondary.feedwater_FWP-1.oil_level")
    ref_j = int(np.nonzero(g["state_names"] == "fw.pump[0].lub.oil_level")[0][0])
    assert np.allclose([float(r[j]) for r in rows[1:]], g["states"][:, ref_j], rtol=1e-9)


def test_step_host_async_equals_step_host():
    """The pipelined host-buffer entry point (pipe_depth launches in flight) produces the same state and outputs as
    the synchronous one, step for step."""
    import torch
    from nuclear_sim_b200 import load_snapshot
    from nuclear_sim_b200 import scenarios as sc
    n, k, steps = 2048, 3, 11
    s0, params = load_snapshot("pwr3000_reactor_dt1")
    pid = np.arange(n)
    st = sc.randomized_states(s0, pid)
    a, b = _sim(st, params), _sim(st, params)
    D = b.pipe_depth
    assert D >= 2
    pin = lambda t: t.pin_memory()
    acts = [pin(torch.from_numpy(sc.load_following_inputs(pid, i * k, k)[0])) for i in range(steps)]
    mags = [pin(torch.from_numpy(sc.load_following_inputs(pid, i * k, k)[1])) for i in range(steps)]
    noise = [pin(torch.from_numpy(sc.noise_inputs(pid, i * k, k))) for i in range(steps)]
    out_a = [(pin(torch.empty((22, n), dtype=torch.float64)), pin(torch.empty(n, dtype=torch.float64)), pin(torch.empty(n, dtype=torch.uint8))) for _ in range(steps)]
    out_b = [(pin(torch.empty((22, n), dtype=torch.float64)), pin(torch.empty(n, dtype=torch.float64)), pin(torch.empty(n, dtype=torch.uint8))) for _ in range(steps)]
    for i in range(steps):
        a.step_host(acts[i], mags[i], noise[i], None, k, *out_a[i])
    tickets = []
    for i in range(steps):
        if i >= D:
            b.wait(tickets[i - D])
        tickets.append(b.step_host_async(acts[i], mags[i], noise[i], None, k, *out_b[i]))
    for t in tickets[-D:]:
        b.wait(t)
    torch.cuda.synchronize()
    for i in range(steps):
        for x, y in zip(out_a[i], out_b[i]):
            assert torch.equal(x, y), f"step {i}"
    assert torch.equal(a.slab, b.slab)


def test_vector_env_and_checkpoint_resume(tmp_path):
    """Vectorised NuclearPlantEnv surface (sim.py:911-940) with in-place reset of scrammed plants, and a checkpoint that
    resumes bit for bit."""
    import torch
    from nuclear_sim_b200 import load_snapshot, field_index
    from nuclear_sim_b200.env import BatchedNuclearPlantEnv, load_checkpoint, save_checkpoint
    s0, params = load_snapshot("pwr3000_reactor_dt1")
    n = 64
    st = np.tile(s0, (n, 1))
    ix = field_index()
    st[::8, ix["pri.fuel_temperature"]] = 1600.0          # every 8th plant scrams on its first step (scram_logic.py:19)
    env = BatchedNuclearPlantEnv(_sim(st, params), auto_reset=True, seed=3)
    assert env.action_space_size == 15 and env.observation_space_size == 22
    obs, rew, done, info = env.step(torch.full((n,), 8))
    assert obs.shape == (n, 22) and rew.shape == (n,) and done.dtype == torch.bool
    assert done.cpu().numpy().tolist() == [(i % 8 == 0) for i in range(n)]
    assert "terminal_observation" in info and info["terminal_observation"].shape == (8, 22)
    # scrammed plants were reset in place (back to the 1600 C initial state), the others kept stepping
    assert float(env.sim.state.scram_status.sum()) == 0.0
    assert (info["episode_steps"].cpu().numpy() == 1).all() and int(env.episode_steps.sum()) == n - 8
    # checkpoint / resume
    env2 = BatchedNuclearPlantEnv(_sim(np.tile(s0, (n, 1)), params), auto_reset=False, seed=5)
    for _ in range(3):
        env2.step(torch.randint(0, 15, (n,), generator=torch.Generator().manual_seed(1)))
    path = str(tmp_path / "ck.pt")
    save_checkpoint(env2.sim, path)
    resumed = load_checkpoint(path)
    acts = torch.full((1, n), 1, dtype=torch.int8)
    a = env2.sim.step(actions=acts, K=4)
    b = resumed.step(actions=acts, K=4)
    assert torch.equal(env2.sim.slab, resumed.slab) and torch.equal(a["observation"], b["observation"])
    # with device-side noise the checkpoint also carries the position of the stream
    s1, p1 = load_snapshot("pwr3000_oil_top_off_dt5")           # constant heat source with noise enabled
    live = _sim(np.tile(s1, (n, 1)), p1)
    live.set_device_rng(11, plant_offset=500)
    live.step(K=3); live.step(K=2)
    save_checkpoint(live, path)
    resumed = load_checkpoint(path)
    live.step(K=4); resumed.step(K=4)
    assert torch.equal(live.slab, resumed.slab)
    assert len(torch.unique(live.state.power_level)) > 1        # the plants really drew different noise


def test_env_cooling_water_reset_cooldowns_and_maintenance_checkpoint(tmp_path):
    """Advisor findings of round 1: env.step(cooling_water_temp=...) (scalar and per plant) reaches the plants; a reset
    plant forgets its threshold cooldown stamps; a checkpoint taken in the middle of a maintenance run resumes with the
    same pending work orders and produces the same orders and state as the uninterrupted run."""
    import json
    import torch
    from nuclear_sim_b200 import load_snapshot, maintenance as M
    from nuclear_sim_b200.env import BatchedNuclearPlantEnv, load_checkpoint, save_checkpoint
    s0, params = load_snapshot("pwr3000_reactor_dt1")
    n = 40
    env = BatchedNuclearPlantEnv(_sim(np.tile(s0, (n, 1)), params), auto_reset=False, seed=1)
    env.step(torch.full((n,), 8), None, 31.5)                      # the reference's positional order: load_demand, cooling water
    assert (env.sim.state["sim.cooling_water_temp"] == 31.5).all()
    cw = torch.linspace(18.0, 30.0, n, dtype=torch.float64)
    obs, *_ = env.step(torch.full((n,), 8), cooling_water_temp=cw, magnitude=torch.ones(n, dtype=torch.float64))
    assert torch.equal(env.sim.state["sim.cooling_water_temp"].cpu(), cw)
    assert torch.allclose(obs[:, 17].cpu(), cw / 35.0)             # observation 17 = cooling water temperature / 35
    # cooldown stamps of a reset plant
    g = np.load(os.path.join(U.GOLDEN, "maint_oil_top_off.npz"), allow_pickle=False)
    log = json.loads(str(g["log"]))
    table = M.ThresholdTable(log["maintenance_system"])
    T = g["states"].shape[0]
    mk = lambda s: M.BatchedAutoMaintenance(s, table, aggressive=True)   # noqa: E731
    full = _sim(np.tile(g["state0"], (3, 1)), g["params"]); m_full = mk(full)
    part = _sim(np.tile(g["state0"], (3, 1)), g["params"]); m_part = mk(part)
    cut = int(log["created"][0]["step"]) + 1                       # a work order is pending at the cut
    m_full.advance(T)
    m_part.advance(cut)
    assert m_part._pending
    path = str(tmp_path / "maint.pt")
    save_checkpoint(part, path, maintenance=m_part)
    resumed, m_res = load_checkpoint(path, maintenance=mk)
    m_res.advance(T - cut)
    key = lambda w: (w.plant, w.work_order_id, w.component_id, w.action, w.created, w.executed_at, w.success)   # noqa: E731
    assert [key(w) for w in m_res.created_log] == [key(w) for w in m_full.created_log]
    assert [key(w) for w in m_res.executed_log] == [key(w) for w in m_full.executed_log] and m_res.executed_log
    assert torch.equal(resumed.slab, full.slab)
    with pytest.raises(ValueError):
        load_checkpoint(path)                                      # stamps without their table would be dropped silently
    fired = (full._thr["last"] > -float("inf")).any(dim=0)
    assert bool(fired.all())
    full.reset(torch.tensor([True, False, False]))
    fired = (full._thr["last"] > -float("inf")).any(dim=0).cpu().numpy().tolist()
    assert fired == [False, True, True]


def test_env_frame_skip_sums_rewards_up_to_done():
    """BatchedNuclearPlantEnv(frame_skip=4): one fused launch per env step; reward = sum of the per-step rewards up to and
    including the scram step, done = scrammed in any of the 4 steps — equal to four frame_skip=1 steps of the same action
    with the same noise."""
    import torch
    from nuclear_sim_b200 import load_snapshot, field_index
    from nuclear_sim_b200.env import BatchedNuclearPlantEnv
    s0, params = load_snapshot("pwr3000_reactor_dt1")
    n = 48
    st = np.tile(s0, (n, 1))
    ix = field_index()
    st[5, ix["pri.coolant_pressure"]] = 17.15            # OPEN valve / rod actions aside, this plant crosses 17.2 MPa within the 4 steps or not at all
    st[9, ix["pri.fuel_temperature"]] = 1600.0           # scrams in the first of the 4 steps (scram_logic.py:19)
    one = BatchedNuclearPlantEnv(_sim(st, params), auto_reset=False, seed=7, frame_skip=1)
    four = BatchedNuclearPlantEnv(_sim(st, params), auto_reset=False, seed=7, frame_skip=4)
    act = torch.full((n,), 1)                            # CONTROL_ROD_WITHDRAW
    # the same draws: frame_skip=4 draws a [4, 5, N] block per step; feed the scalar env the rows of that block
    gen = torch.Generator(device="cpu").manual_seed(7)
    z = torch.randn((4, 2, n), generator=gen, dtype=torch.float64); u = torch.rand((4, 3, n), generator=gen, dtype=torch.float64)
    block = torch.cat([z, u], dim=1)
    rew = torch.zeros(n, dtype=torch.float64, device="cuda:0"); done_any = torch.zeros(n, dtype=torch.bool, device="cuda:0")
    for k in range(4):
        out = one.sim.step(actions=act.to(torch.int8).reshape(1, -1), noise=block[k:k + 1], K=1)
        rew += out["reward"] * (~done_any)
        done_any |= out["done"]
    obs4, rew4, done4, info4 = four.step(act)
    assert torch.equal(four.sim.slab, one.sim.slab)
    assert torch.equal(done4, done_any) and bool(done4[9]) and int(done4.sum()) >= 1
    assert torch.allclose(rew4, rew, rtol=1e-13, atol=0.0)
    assert int(info4["episode_steps"][0]) == 4


def test_cuda_batched_timing_sweep():
    """The batched TimingOptimizer replacement on the CUDA engine: 512 candidate oil levels in one batch; the committed
    initial level reproduces the reference's trigger time and the curve is monotone."""
    import json
    import os
    from nuclear_sim_b200 import optimize as O, field_index
    g = np.load(os.path.join(U.GOLDEN, "maint_oil_top_off.npz"), allow_pickle=False)
    log = json.loads(str(g["log"]))
    fld = "fw.pump[0].lub.oil_level"
    level0 = float(g["state0"][field_index()[fld]])
    vals = np.sort(np.append(np.linspace(58.2, 61.5, 511), level0))
    hrs = O.trigger_time_sweep(g["state0"], g["params"], log["maintenance_system"], fld, vals, "oil_top_off", 5.0, component_id="FWP-1")
    assert hrs[int(np.nonzero(vals == level0)[0][0])] == log["created"][0]["t"] / 60.0
    assert not np.isnan(hrs).any() and np.all(np.diff(hrs) >= 0) and hrs[-1] > hrs[0]


def test_c_abi_from_plain_c(tmp_path):
    """examples/step_from_c.c (gcc, no Python / torch on its side) steps 96 identical plants through the C ABI; the state,
    observation and reward it writes are the Python host API's, bit for bit."""
    import subprocess
    from nuclear_sim_b200 import load_snapshot
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "step_from_c")
    lib_dir = os.path.join(root, "nuclear-sim_b200", "_lib")
    subprocess.check_call(["gcc", "-O2", "-I", os.path.join(root, "include"), "-I", "/usr/local/cuda/include",
                           os.path.join(root, "examples", "step_from_c.c"), "-o", exe, "-L", lib_dir, "-lnps_b200",
                           "-L", "/usr/local/cuda/lib64", "-lcudart", f"-Wl,-rpath,{lib_dir}"])
    s0, params = load_snapshot("pwr3000_reactor_dt1")
    np.concatenate([s0, params]).astype(np.float64).tofile(str(tmp_path / "plant.bin"))
    n, launches, k = 96, 3, 5
    out = subprocess.run([exe, str(tmp_path / "plant.bin"), str(n), str(launches), str(k), str(tmp_path / "out.bin")],
                         capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    got = np.fromfile(str(tmp_path / "out.bin"))
    sim = _sim(np.tile(s0, (n, 1)), params)
    for _ in range(launches):
        res = sim.step(K=k)
    want = np.concatenate([sim.state_numpy()[0], res["observation"][0].cpu().numpy(), [float(res["reward"][0])]])
    np.testing.assert_array_equal(got.view(np.uint64), want.view(np.uint64))


def test_error_conventions_through_the_abi():
    """Return codes and nps_last_error(): nothing fails silently (include/nps_b200.h conventions)."""
    import torch
    from nuclear_sim_b200 import _clib, load_snapshot
    L = _clib.lib()
    s0, params = load_snapshot("pwr3000_reactor_dt1")
    sim = _sim(np.tile(s0, (4, 1)), params)
    h = ctypes.c_void_p()
    assert L.nps_create(ctypes.c_int64(0), 0, ctypes.byref(h)) < 0 and b"bad arguments" in L.nps_last_error()
    assert L.nps_create(ctypes.c_int64(4), 99, ctypes.byref(h)) < 0 and b"device index" in L.nps_last_error()
    with pytest.raises(_clib.NpsError, match="k_substeps"):
        sim.step(K=0)
    with pytest.raises(_clib.NpsError, match="set_thresholds"):
        sim.check_thresholds()
    ring = torch.zeros(8, dtype=torch.float64, device="cuda")
    assert L.nps_log_row(sim._h, ctypes.c_void_p(sim.slab.data_ptr()), ctypes.c_void_p(ring.data_ptr()), ctypes.c_int64(1),
                         ctypes.c_int64(0), None) < 0 and b"no logged fields" in L.nps_last_error()
    last = torch.zeros(4, dtype=torch.float64, device="cuda")
    flags = torch.zeros(4, dtype=torch.int32, device="cuda")
    assert L.nps_check_thresholds(sim._h, ctypes.c_void_p(sim.slab.data_ptr()), ctypes.c_void_p(last.data_ptr()),
                                  ctypes.c_void_p(flags.data_ptr()), None, None) < 0 and b"no thresholds" in L.nps_last_error()
    bad = np.array([10 ** 6], dtype=np.int32)
    assert L.nps_set_logged_fields(sim._h, bad.ctypes.data_as(ctypes.c_void_p), 1) < 0 and b"out of range" in L.nps_last_error()
    assert L.nps_wait(sim._h, 17) < 0 and b"bad ticket" in L.nps_last_error()
    # every target has a restatement now; an action a target does not know fails the way the reference's does
    # (success False, no state change): oil_change on the SG system object
    import torch
    from nuclear_sim_b200 import struct_range
    before = sim.slab.clone()
    assert sim.apply_maintenance([(0, 8, 0, 0)]) == [0]
    lo, _ = struct_range("rep")             # the report columns are refreshed after every maintenance call
    assert torch.equal(before[:lo], sim.slab[:lo])


def test_device_pow_against_libm():
    """The step kernel's power function (exact specialisations, fastpow.h, libdevice fallback) against numpy's pow
    (glibc, < 1 ulp) on the operand classes of the model and on every special case: <= 2 ulp everywhere it is finite,
    identical for exact cases and specials."""
    import torch
    from nuclear_sim_b200 import _clib
    L = _clib.lib()
    rng = np.random.RandomState(5)
    n = 400000
    x = 10.0 ** rng.uniform(-6, 6, n)
    y = rng.uniform(-3, 4, n)
    x[::7] = 1.0 + rng.uniform(-1e-3, 1e-3, len(x[::7]))                    # near 1
    y[::11] = rng.choice([0.38, 0.8, 0.15, -0.6, 1.8, 2.2, 2.4, 1.6, 1.4, 1.3, 1.0 / 3, 2.0, 0.5, 3.0, 1.5, 0.25, 1.0, 0.0], len(y[::11]))
    sx = np.array([0.0, -0.0, -2.0, -2.0, np.inf, np.nan, 5e-324, 1e-310, 2.0, 2.0, 1e300, 1e-300, 0.5, 1e150, 3.0])
    sy = np.array([1.8, 1.8, 3.0, 0.5, 0.8, 1.2, 0.7, 0.7, np.inf, np.nan, 1.5, 1.5, 1100.0, 2.5, -2.0])
    x, y = np.concatenate([x, sx]), np.concatenate([y, sy])
    dx, dy = torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda()
    out = torch.empty_like(dx)
    assert L.nps_selftest_pow(ctypes.c_void_p(dx.data_ptr()), ctypes.c_void_p(dy.data_ptr()), ctypes.c_void_p(out.data_ptr()),
                              ctypes.c_int64(len(x)), None) == 0
    torch.cuda.synchronize()
    got = out.cpu().numpy()
    with np.errstate(all="ignore"):
        want = np.power(x, y)
    fin = np.isfinite(want) & (want != 0)
    ulps = np.abs(got[fin].view(np.int64) - want[fin].view(np.int64))
    assert ulps.max() <= 2, (ulps.max(), x[fin][ulps.argmax()], y[fin][ulps.argmax()])
    assert (ulps > 1).mean() < 1e-3
    rest = ~fin
    np.testing.assert_array_equal(np.isnan(got[rest]), np.isnan(want[rest]))
    ok = ~np.isnan(want[rest])
    np.testing.assert_array_equal(got[rest][ok], want[rest][ok])
    # the exact specialisations are what they say they are: x*x (the correctly rounded square), x, 1
    np.testing.assert_array_equal(got[y == 2.0], (x * x)[y == 2.0])
    np.testing.assert_array_equal(got[(y == 1.0) & ~np.isnan(x)], x[(y == 1.0) & ~np.isnan(x)])
    np.testing.assert_array_equal(got[(y == 0.0) & ~np.isnan(x)], 1.0)


def test_batched_export_trajectory_in_reference_schema(tmp_path):
    """Device ring buffer -> the reference's wide CSV for one plant of a batch: header and every value equal what the
    host row store writes from the full states of the same steps."""
    import datetime as dt
    from nuclear_sim_b200 import load_snapshot
    from nuclear_sim_b200.export import TrajectoryStore
    from nuclear_sim_b200 import scenarios as sc
    s0, params = load_snapshot("pwr3000_oil_top_off_dt5")
    n = 10
    sim = _sim(sc.randomized_states(s0, np.arange(n)), params)
    n_fields = sim.set_logged_columns("secondary.steam_generator_SG-1.", ring_rows=8)
    assert n_fields > 20
    store = TrajectoryStore(dt.datetime(2024, 1, 1))
    for k in range(6):
        sim.step()
        sim.log_row()
        store.add_row(dt.datetime(2024, 1, 1) + dt.timedelta(minutes=5.0 * (k + 1)), sim.state_numpy()[7])
    assert sim.export_trajectory(7, str(tmp_path / "ring.csv")) == 6
    store.export_by_subcategory("secondary", "steam_generator_SG-1", str(tmp_path / "store.csv"))
    assert open(tmp_path / "ring.csv").read() == open(tmp_path / "store.csv").read()


def test_device_export_equals_the_reference_state_log(tmp_path):
    """tests/golden/ref_state_log_oil_top_off.csv is the reference's own StateManager.export_to_csv of the first 12 steps
    of the maint_oil_top_off scenario.  The device path — step kernel, TMA ring-buffer rows of every exportable field,
    export_trajectory for one plant of a batch — must write the same header and, row by row, the same 788 values
    (numbers to 1e-9, status strings and booleans exactly)."""
    import csv
    import torch
    g = np.load(os.path.join(U.GOLDEN, "maint_oil_top_off.npz"), allow_pickle=False)
    ref = list(csv.reader(open(os.path.join(U.GOLDEN, "ref_state_log_oil_top_off.csv"), newline="")))
    T = len(ref) - 1
    n, me = 6, 4
    rng = np.random.RandomState(9)
    st = np.tile(g["state0"], (n, 1)) * (1.0 + 1e-3 * rng.standard_normal((n, 1)))
    st[me] = g["state0"]
    sim = _sim(st, g["params"])
    assert sim.set_logged_columns(None, ring_rows=16) > 600
    for t in range(T):
        z = np.tile(g["noise"][t][None, :, None], (1, 1, n))          # [1, 5, n]
        sim.step(noise=torch.from_numpy(np.ascontiguousarray(z)), K=1)
        sim.log_row()
    assert sim.export_trajectory(me, str(tmp_path / "ours.csv")) == T
    ours = list(csv.reader(open(tmp_path / "ours.csv", newline="")))
    assert ours[0] == ref[0] and len(ours[0]) == 789
    bad = []
    for r in range(1, T + 1):
        for j in range(1, 789):
            a, b = ours[r][j], ref[r][j]
            if a == b:
                continue
            try:
                fa, fb = float(a), float(b)
            except ValueError:
                bad.append((r, ref[0][j], a, b))
                continue
            if not abs(fa - fb) <= 1e-9 * max(1e-6, abs(fb)) * r:
                bad.append((r, ref[0][j], a, b))
    assert not bad, f"{len(bad)} cells differ from the reference's state log, e.g. {bad[:5]}"


# ---- in-launch monitoring: the step of every discrete event does not depend on how steps are fused into launches ---
def _monitored_run(g, kmax, table_cfg=None):
    """Replay a trajectory fixture with launches of up to kmax substeps and the in-launch monitor; returns the simulator,
    the drained threshold events and the per-step done flags."""
    import torch
    from nuclear_sim_b200 import maintenance as M
    sim = _sim(g["state0"], g["params"])
    if table_cfg is not None:
        sim.set_thresholds(M.ThresholdTable(table_cfg).device_rows())
    sim.enable_monitor(per_substep=True, max_k=kmax)
    T = g["actions"].shape[0]
    inj = g["inject"]
    done = np.zeros((T, sim.n_plants), dtype=bool)
    rewards = np.zeros((T, sim.n_plants))
    events = []
    t = 0
    while t < T:
        for p in range(sim.n_plants):
            if not np.isnan(inj[t, p, 0]):
                sim.slab[int(inj[t, p, 0]), p] = float(inj[t, p, 1])
        k = 1
        while t + k < T and k < kmax and np.isnan(inj[t + k, :, 0]).all():
            k += 1
        out = sim.step(actions=torch.from_numpy(np.ascontiguousarray(g["actions"][t:t + k])),
                       magnitudes=torch.from_numpy(np.ascontiguousarray(g["magnitudes"][t:t + k])),
                       noise=torch.from_numpy(np.ascontiguousarray(g["noise"][t:t + k].transpose(0, 2, 1))),
                       power_setpoint=torch.from_numpy(np.ascontiguousarray(g["setpoint"][t:t + k])), K=k)
        done[t:t + k] = out["done_k"].cpu().numpy()
        rewards[t:t + k] = out["reward_k"].cpu().numpy()
        events.append(sim.drain_step_events())
        t += k
    return sim, np.concatenate(events), done, rewards


@pytest.mark.parametrize("name", ["cfg4_scram", "cfg6_secondary_trips", "cfg7_turbine_trips_fouling"])
def test_fused_launches_report_the_same_event_steps_as_single_steps(name):
    """K = 64 against K = 1: first scram step, first step of every watched trip latch, threshold violations
    (plant, row, step, value, time), per-step done and reward, status words and the final state — all identical;
    scram steps also equal the live-reference fixture."""
    import json
    import torch
    g = U.load_golden(name)
    cfg = json.loads(str(np.load(os.path.join(U.GOLDEN, "maint_oil_top_off.npz"), allow_pickle=False)["log"]))["maintenance_system"]
    a, ev_a, done_a, rew_a = _monitored_run(g, 1, cfg)
    b, ev_b, done_b, rew_b = _monitored_run(g, 64, cfg)
    assert a.n_launches > b.n_launches * 8
    assert torch.equal(a.slab, b.slab)
    assert torch.equal(a.first_scram_step, b.first_scram_step) and torch.equal(a.status, b.status)
    assert torch.equal(a.first_nan_reset_step, b.first_nan_reset_step)
    for (wa, sa), (wb, sb) in zip(a.watch_steps().items(), b.watch_steps().items()):
        assert wa == wb and torch.equal(sa, sb), wa
    assert ev_a.tolist() == ev_b.tolist()
    assert np.array_equal(done_a, done_b) and np.array_equal(rew_a, rew_b)
    first_done = np.where(done_a.any(0), done_a.argmax(0), -1)
    assert first_done.tolist() == g["done_step"].tolist()
    assert a.first_scram_step.cpu().numpy().tolist() == g["done_step"].tolist()
    if name != "cfg4_scram":
        assert len(ev_a) > 0                                        # degraded plants do cross maintenance thresholds
        trips = torch.stack(list(a.watch_steps().values()))
        assert int((trips >= 0).sum()) >= 2                         # and do latch trips, at known steps


@pytest.mark.parametrize("name", [t for t in TRIP_SCENARIOS if t not in ("trip_vacuum_lag_rotation", "trip_cond_tube_vibration",
                                                                       "trip_sg_no_load_balancing", "trip_fw_manual_flow",
                                                                       "trip_rotor_slow", "trip_ejector_out_of_range")])
def test_trip_latch_steps_equal_the_reference_inside_one_fused_launch(name):
    """Each trip_* fixture latches one protection path of the live reference.  All 60 steps run as ONE monitored launch;
    the step the monitor stamps for every watched latch must be the step at which the reference's flag first reads 1
    (the fixtures checkpoint steps 1..6 one by one), and latches the reference never sets must stay unstamped."""
    import torch
    from nuclear_sim_b200 import field_index
    g = U.load_golden(name)
    ix = field_index()
    watch = ["turb.prot_trip_active", "fw.prot_system_trip_active", "cond.vs_trip_high_pressure", "cond.vs_alarm_high_pressure",
             "turb.vib_displacement_alarm", "turb.vib_critical_speed_alarm", "fw.prot_npsh_low_alarm_active",
             "fw.prot_npsh_critical_trip_active"]
    sim = _sim(g["state0"], g["params"])
    sim.enable_monitor(watch=watch, max_k=60)
    _advance(sim, g, 0, 60, kmax=60)
    assert sim.n_launches == 1
    got = {w: int(s[0]) for w, s in sim.watch_steps().items()}
    cps = [int(c) for c in g["checkpoints"]]
    n_latched = 0
    for w in watch:
        col = g["states"][:, 0, ix[w]]
        if float(g["state0"][0, ix[w]]) != 0.0:
            continue
        on = np.flatnonzero(col != 0.0)
        if len(on) == 0:
            assert got[w] == -1, (w, got[w])
            continue
        c = int(on[0])
        lo = cps[c - 1] if c > 0 else 0              # flag still 0 after `lo` steps, 1 after cps[c] steps
        assert lo <= got[w] < cps[c], (w, got[w], lo, cps[c])
        n_latched += 1
    assert n_latched >= 1, "the fixture latches nothing that is watched"
    U.assert_states_close(sim.state_numpy(), g["states"][-1], U.TOL_STEP * 60, f"{name} step 60 (one launch)")


@pytest.mark.parametrize("driver", ["advance_interleaved", "advance_threaded"])
def test_interleaved_parts_on_streams_equal_one_batch(driver):
    """advance_interleaved (one host thread resuming the batches in turn) / advance_threaded (one host thread per batch) on the device: 1 024 plants as one batch against the same plants as two batches of 512 with their
    own CUDA streams, resumed in turn - same work orders at the same times, bit-identical final states."""
    import json
    import torch
    from nuclear_sim_b200 import maintenance as M, field_index
    g = np.load(os.path.join(U.GOLDEN, "maint_oil_top_off.npz"), allow_pickle=False)
    cfg = json.loads(str(g["log"]))["maintenance_system"]
    ix = field_index()
    n = 1024
    st = np.tile(g["state0"], (n, 1))
    rng = np.random.RandomState(8)
    st[:, ix["fw.pump[0].lub.oil_level"]] = 58.0 + rng.uniform(-0.3, 1.5, n)
    st[:, ix["fw.pump[2].lub.oil_level"]] = 58.0 + rng.uniform(0.0, 3.0, n)
    st[:, ix["fw.pump[1].lub.oil_contamination_level"]] = 15.2 - rng.uniform(-0.01, 0.03, n)
    whole_sim = _sim(st, g["params"])
    whole = M.NativeAutoMaintenance(whole_sim, M.ThresholdTable(cfg), aggressive=True)
    whole.advance(30)
    part_sims = [_sim(st[:512], g["params"]), _sim(st[512:], g["params"])]
    parts = [M.NativeAutoMaintenance(q, M.ThresholdTable(cfg), aggressive=True) for q in part_sims]
    getattr(M, driver)(parts, 30)
    torch.cuda.synchronize()
    key = lambda w, off: (w.created, w.plant + off, w.component_id, w.action, w.work_order_id,  # noqa: E731
                          -1.0 if w.executed_at is None else w.executed_at, bool(w.success))
    for m in [whole] + parts:
        m.materialize_logs()
    a = sorted(key(w, 0) for w in whole.created_log)
    b = sorted([key(w, 0) for w in parts[0].created_log] + [key(w, 512) for w in parts[1].created_log])
    assert a == b and len(a) > 500
    assert torch.equal(torch.cat([q.slab for q in part_sims], dim=1), whole_sim.slab)


def test_status_word_reports_the_nan_reset():
    """thermal_hydraulics.py:257-269 resets five primary fields when one of them is NaN — silently in the reference; the
    batched engine reproduces the reset and raises bit 0 of the plant's status word, with the step it happened at."""
    import torch
    from nuclear_sim_b200 import field_index, load_snapshot
    s0, params = load_snapshot("pwr3000_reactor_dt1")
    n = 70
    sim = _sim(np.tile(s0, (n, 1)), params)
    sim.enable_monitor()
    sim.step(K=3)
    sim.state["pri.fuel_temperature"][5] = float("nan")
    sim.state["pri.coolant_pressure"][64] = float("nan")
    sim.step(K=4)
    st = sim.status.cpu().numpy()
    fr = sim.first_nan_reset_step.cpu().numpy()
    assert [int(p) for p in np.nonzero(st & 1)[0]] == [5, 64]
    assert fr[5] == 3 and fr[64] == 3 and (np.delete(fr, [5, 64]) == -1).all()
    assert float(sim.state.power_level[5]) > 0 and not torch.isnan(sim.slab[:, 5]).any()
    sim.reset()
    assert int(sim.status.sum()) == 0 and int((sim.first_nan_reset_step >= 0).sum()) == 0


@pytest.mark.parametrize("name", MAINT_SCENARIOS)
def test_cuda_maintenance_advance_with_fused_launches(name):
    """BatchedAutoMaintenance.advance on the device: in-launch threshold evaluation, launches cut at the 15-minute gate
    steps — threshold events, work orders created / executed and the final state equal the live-reference fixture."""
    import json
    import torch
    from nuclear_sim_b200 import maintenance as M
    from tests.test_maintenance_host import compare_logs
    g = np.load(os.path.join(U.GOLDEN, f"maint_{name}.npz"), allow_pickle=False)
    log = json.loads(str(g["log"]))
    T = g["states"].shape[0]
    sim = _sim(g["state0"][None, :], g["params"])
    maint = M.BatchedAutoMaintenance(sim, M.ThresholdTable(log["maintenance_system"]), aggressive=True)
    noise = torch.from_numpy(np.ascontiguousarray((g["noise"][:, None, :] if g["noise"].ndim == 2 else g["noise"]).transpose(0, 2, 1)))
    maint.advance(T, noise=noise, max_k=64)
    assert sim.n_launches < 3 * T      # step + flag kernel per gate instead of per step (+ a few maintenance launches)
    compare_logs(maint, log)
    U.assert_states_close(sim.state_numpy(), g["states"][T - 1][None, :], U.TOL_STEP * T, f"{name} final state")


# ---- BASELINE sizes against the live reference: 64 reference plants embedded in the full batch -----------------------
def _run_sized_embedded(name, kmax):
    import torch
    from nuclear_sim_b200 import load_snapshot
    from nuclear_sim_b200 import scenarios as sc
    g = U.load_sized(name)
    n = int(g["total"])
    ids = g["plant_ids"].astype(np.int64)
    s0, params = load_snapshot("pwr3000_reactor_dt1")
    assert np.array_equal(params, g["params"])
    pid = np.arange(n)
    st = sc.randomized_states(s0, pid)          # the bench workload's plants ...
    st[ids] = g["state0"]                       # ... with the reference-built plants at their scattered global ids
    sim = _sim(st, g["params"])
    sim.enable_monitor(per_substep=True, max_k=kmax)
    T = int(g["n_steps"])
    inj = g["inject"]
    inj_steps = sorted(set(np.nonzero(~np.isnan(inj[:, :, 0, 0]))[0].tolist()))
    cps = [int(c) for c in g["checkpoints"]]
    stops = sorted(set(inj_steps + cps + [T]))
    dev_ids = torch.as_tensor(ids, device="cuda:0")
    done = np.zeros((T, len(ids)), dtype=bool)
    t = 0
    while t < T:
        for j in np.nonzero(~np.isnan(inj[t, :, 0, 0]))[0]:
            for f, v in inj[t, j]:
                if not np.isnan(f):
                    sim.slab[int(f), int(ids[j])] = float(v)
        nxt = min(s for s in stops if s > t)
        k = min(kmax, nxt - t)
        acts, mags = sc.load_following_inputs(pid, t, k)
        acts[:, ids] = g["actions"][t:t + k]
        mags[:, ids] = g["magnitudes"][t:t + k]
        noise = torch.zeros((k, 5, n), dtype=torch.float64, device="cuda:0")
        noise[:, 2:, :] = 1.0
        noise[:, :, dev_ids] = torch.from_numpy(np.ascontiguousarray(g["noise"][t:t + k].transpose(0, 2, 1))).to("cuda:0")
        out = sim.step(actions=torch.from_numpy(acts), magnitudes=torch.from_numpy(mags), noise=noise, K=k)
        done[t:t + k] = out["done_k"][:, dev_ids].cpu().numpy()
        t += k
        if t in cps:
            got = sim.slab[:, dev_ids].t().contiguous().cpu().numpy()
            tol = U.TOL_STEP * max(1, min(t, 1000)) if t < 3600 else U.TOL_LONG
            U.assert_states_close(got, g["states"][cps.index(t)], tol, f"{name} step {t}")
    return sim, g, done, dev_ids


def test_cfg3_reference_plants_embedded_in_65536():
    """BASELINE config #3 at full size: 65 536 plants, 3 600 steps, 64-substep launches; the 64 plants built by the
    reference's randomised-IC generators and stepped by the reference must come out of the full batch as the fixture
    recorded them (1e-9 per step, 1e-6 after 3 600 steps), whatever their neighbours do."""
    sim, g, done, dev_ids = _run_sized_embedded("cfg3_rand64", 64)
    assert not done.any() and (sim.first_scram_step[dev_ids] < 0).all()
    assert float(sim.state.power_level.mean()) >= 0.0


def test_cfg4_reference_transients_embedded_in_16384():
    """BASELINE config #4 at full size: 16 384 plants, 900 steps; 64 plants get the reference's EquipmentFailureSimulator
    / emergency transients at their own trigger steps.  done every step, first scram step, and the states at the
    checkpoints equal the live reference's."""
    sim, g, done, dev_ids = _run_sized_embedded("cfg4_fail64", 64)
    assert np.array_equal(done, g["done"].astype(bool))
    assert sim.first_scram_step[dev_ids].cpu().numpy().tolist() == g["first_scram_step"].tolist()
    assert (g["first_scram_step"] >= 0).sum() >= 15


@pytest.mark.parametrize("name", ["cfg3_loadfollow", "cfg6_secondary_trips", "cfg7_turbine_trips_fouling", "cfg1_oil_top_off"])
def test_split_launch_shape_is_bit_identical_to_one_thread_per_plant(name):
    """Small batches run two threads per plant (source half / sink half, pipelined by one substep).  State, observation,
    reward, done, per-substep reward / done, threshold events, watch stamps must equal the one-thread-per-plant shape
    bit for bit, with and without monitoring, for K = 1 and fused launches."""
    import json
    import torch
    from nuclear_sim_b200 import maintenance as M
    g = U.load_golden(name)
    cfg = json.loads(str(np.load(os.path.join(U.GOLDEN, "maint_oil_top_off.npz"), allow_pickle=False)["log"]))["maintenance_system"]
    T = min(120, g["actions"].shape[0])
    sims = []
    for shape in (0, 1):
        sim = _sim(g["state0"], g["params"])
        sim.set_small_batch_shape(shape)
        sim.set_thresholds(M.ThresholdTable(cfg).device_rows())
        sim.enable_monitor(per_substep=True, max_k=32)
        sims.append(sim)
    outs = [[], []]
    t = 0
    for k in [1, 2, 5, 32, 17, 1, 32, 30]:
        k = min(k, T - t)
        if k <= 0:
            break
        for i, sim in enumerate(sims):
            o = sim.step(actions=torch.from_numpy(np.ascontiguousarray(g["actions"][t:t + k])),
                         magnitudes=torch.from_numpy(np.ascontiguousarray(g["magnitudes"][t:t + k])),
                         noise=torch.from_numpy(np.ascontiguousarray(g["noise"][t:t + k].transpose(0, 2, 1))),
                         power_setpoint=torch.from_numpy(np.ascontiguousarray(g["setpoint"][t:t + k])), K=k)
            outs[i].append({q: v.clone() for q, v in o.items()})
        t += k
        assert torch.equal(sims[0].slab, sims[1].slab), f"state differs after step {t}"
    for a, b in zip(*outs):
        for q in a:
            assert torch.equal(a[q], b[q]) or (torch.isnan(a[q]) == torch.isnan(b[q])).all() and torch.equal(torch.nan_to_num(a[q]), torch.nan_to_num(b[q])), q
    ev = [s.drain_step_events() for s in sims]
    assert ev[0].tolist() == ev[1].tolist()
    for (wa, sa), (wb, sb) in zip(sims[0].watch_steps().items(), sims[1].watch_steps().items()):
        assert torch.equal(sa, sb), wa
    assert torch.equal(sims[0].first_scram_step, sims[1].first_scram_step) and torch.equal(sims[0].status, sims[1].status)


@pytest.mark.parametrize("n,k", [(4099, 6), (18913, 3)])
def test_split_launch_shape_on_a_ragged_random_batch(oracle_lib, n, k):
    """4 099 and 18 913 plants (ragged last block; the second is just under the split shape's one-wave limit of 18 944),
    random actions and noise, 3 fused launches: split shape == one thread per plant bitwise, and both within tolerance
    of the host oracle."""
    import torch
    from nuclear_sim_b200 import load_snapshot
    from nuclear_sim_b200 import scenarios as sc
    s0, params = load_snapshot("pwr3000_reactor_dt1")
    pid = np.arange(n)
    st = sc.randomized_states(s0, pid)
    a, b = _sim(st, params), _sim(st, params)
    b.set_small_batch_shape(1)
    ref = st.copy()
    for i in range(3):
        acts, mags = sc.load_following_inputs(pid, i * k, k)
        noise = sc.noise_inputs(pid, i * k, k)
        for sim in (a, b):
            sim.step(actions=torch.from_numpy(acts), magnitudes=torch.from_numpy(mags), noise=torch.from_numpy(noise), K=k)
        ref = U.oracle_run(oracle_lib, ref, params, acts, mags, np.ascontiguousarray(noise.transpose(0, 2, 1)),
                           np.full((k, n), np.nan), None, 0, k)
    assert torch.equal(a.slab, b.slab)
    U.assert_states_close(a.state_numpy(), ref, U.TOL_STEP * 3 * k, "split shape vs oracle")
