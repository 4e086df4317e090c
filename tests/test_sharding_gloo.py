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
