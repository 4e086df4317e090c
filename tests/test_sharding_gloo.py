"""CPU, world_size 2 over gloo: the multi-rank host path — contiguous plant sharding, shard-invariant
initial conditions / inputs, and the end-of-run all_gather of trajectory summaries."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_total, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from nuclear_sim_b200 import load_snapshot
    from nuclear_sim_b200 import scenarios as sc
    s0, _ = load_snapshot("pwr3000_reactor_dt1")
    n = n_total // world
    pid = np.arange(rank * n, (rank + 1) * n)
    st = sc.randomized_states(s0, pid)
    acts, mags = sc.load_following_inputs(pid, 16, 8)
    summary = torch.from_numpy(np.stack([st.sum(axis=1), acts.astype(np.float64).sum(axis=0), mags.sum(axis=0)], axis=1))
    gathered = [torch.empty_like(summary) for _ in range(world)]
    dist.all_gather(gathered, summary)
    t = torch.tensor([float(rank + 1)])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        np.save(os.path.join(out_dir, "gathered.npy"), torch.cat(gathered).numpy())
        np.save(os.path.join(out_dir, "tmax.npy"), t.numpy())
    dist.destroy_process_group()


def test_sharded_inputs_match_single_process(tmp_path):
    from nuclear_sim_b200 import load_snapshot
    from nuclear_sim_b200 import scenarios as sc
    n_total, world = 64, 2
    mp.spawn(_worker, args=(world, _free_port(), n_total, str(tmp_path)), nprocs=world, join=True)
    got = np.load(tmp_path / "gathered.npy")
    s0, _ = load_snapshot("pwr3000_reactor_dt1")
    pid = np.arange(n_total)
    st = sc.randomized_states(s0, pid)
    acts, mags = sc.load_following_inputs(pid, 16, 8)
    ref = np.stack([st.sum(axis=1), acts.astype(np.float64).sum(axis=0), mags.sum(axis=0)], axis=1)
    np.testing.assert_array_equal(got, ref)
    assert np.load(tmp_path / "tmax.npy")[0] == world   # max-over-ranks reduction used for timing


def test_randomized_states_are_plant_local():
    from nuclear_sim_b200 import load_snapshot
    from nuclear_sim_b200 import scenarios as sc
    s0, _ = load_snapshot("pwr3000_reactor_dt1")
    a = sc.randomized_states(s0, np.arange(100))
    b = sc.randomized_states(s0, np.arange(50, 60))
    np.testing.assert_array_equal(a[50:60], b)
    ids = sc.ic_field_ids()
    assert len(ids) > 100
    changed = (a != s0[None, :]).any(axis=0)
    assert changed[ids].mean() > 0.2 and not changed[np.setdiff1d(np.arange(len(s0)), ids)].any()


def _step_worker(rank, world, port, n_total, out_dir):
    """Each rank steps ITS shard (host oracle engine) with inputs derived from global plant ids, then all-gathers."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from nuclear_sim_b200 import load_snapshot
    from nuclear_sim_b200.sharded import ShardedBatchedSimulator
    from tests import _util as U
    s0, params = load_snapshot("pwr3000_oil_top_off_dt5")       # constant heat source with noise: the draws matter
    sh = ShardedBatchedSimulator(n_total, s0, params, engine_factory=lambda st, p, dev: U.OracleSim(st, p))
    assert (sh.rank, sh.world) == (rank, world)
    for _ in range(2):
        acts, mags = sh.load_following_inputs(3)
        sh.step(actions=acts, magnitudes=mags, noise=sh.noise_inputs(3), K=3)
    full = sh.gather_states()
    summ = sh.gather_summaries(["pri.power_level", "sec.electrical_power_output", "sim.time_minutes"])
    if rank == 0:
        np.save(os.path.join(out_dir, "states.npy"), full.numpy())
        np.save(os.path.join(out_dir, "summary.npy"), summ.numpy())
    dist.destroy_process_group()


def test_two_stepping_shards_equal_one_process(tmp_path):
    """ShardedBatchedSimulator over 2 gloo ranks (ragged: 33 plants -> 17 + 16): the gathered states equal a
    single-process run of the whole batch bit for bit — ICs, actions and noise are functions of the global plant id."""
    from nuclear_sim_b200 import field_index, load_snapshot
    from nuclear_sim_b200.sharded import ShardedBatchedSimulator, shard_range
    from tests import _util as U
    n_total, world = 33, 2
    assert [shard_range(n_total, r, world) for r in range(world)] == [(0, 17), (17, 33)]
    mp.spawn(_step_worker, args=(world, _free_port(), n_total, str(tmp_path)), nprocs=world, join=True)
    s0, params = load_snapshot("pwr3000_oil_top_off_dt5")
    one = ShardedBatchedSimulator(n_total, s0, params, rank=0, world=1, engine_factory=lambda st, p, dev: U.OracleSim(st, p))
    for _ in range(2):
        acts, mags = one.load_following_inputs(3)
        one.step(actions=acts, magnitudes=mags, noise=one.noise_inputs(3), K=3)
    ref = one.sim.state_numpy()
    got = np.load(tmp_path / "states.npy")
    assert got.shape == ref.shape
    np.testing.assert_array_equal(got.view(np.uint64), ref.view(np.uint64))
    ix = field_index()
    summ = np.load(tmp_path / "summary.npy")
    np.testing.assert_array_equal(summ[:, 0], ref[:, ix["pri.power_level"]])
    assert len(np.unique(ref[:, ix["pri.hs_raw_noise_mw"]])) == n_total      # per-plant noise really differs
