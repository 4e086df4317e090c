"""Vectorised RL environment and checkpointing on top of the batched engine.

``BatchedNuclearPlantEnv`` mirrors the reference's gym-style wrapper ``NuclearPlantEnv``
(nuclear_simulator/simulator/core/sim.py:911-940: ``step(action_idx, load_demand, cooling_water_temp) ->
(observation, reward, done, info)``, ``reset() -> observation``, ``action_space_size`` = 15 ControlAction values,
``observation_space_size`` = 22) with a leading plant axis; plants that scram (``done``) can be reset in place on the
device without touching the others (``auto_reset``).

``save_checkpoint`` / ``load_checkpoint``: the reference has no checkpoint/resume; here a plant batch is one
tensor plus a parameter vector, so a checkpoint is a ``torch.save`` of exactly that and resuming is bit-exact.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch


N_ACTIONS = 15          # len(ControlAction): systems/primary/__init__.py:28-45


class BatchedNuclearPlantEnv:
    def __init__(self, sim, auto_reset: bool = True, seed: int = 0, frame_skip: int = 1):
        """frame_skip = K repeats each action for K reference steps in ONE fused launch: the reward is the sum of the
        per-step rewards up to and including the step a plant scrammed at, done is "scrammed in any of the K steps"
        (the in-launch monitor supplies reward and done of every substep), the observation is the last step's."""
        self.sim = sim
        self.frame_skip = int(frame_skip)
        if self.frame_skip > 1:
            sim.enable_monitor(per_substep=True, max_k=self.frame_skip)
        self.n_plants = sim.n_plants
        self.action_space_size = N_ACTIONS
        self.observation_space_size = 22
        self.auto_reset = auto_reset
        self._gen = torch.Generator(device="cpu").manual_seed(seed)
        self.episode_steps = torch.zeros(self.n_plants, dtype=torch.int64, device=sim.device)

    def _noise(self) -> torch.Tensor:
        k = self.frame_skip
        z = torch.randn((k, 2, self.n_plants), generator=self._gen, dtype=torch.float64)
        u = torch.rand((k, 3, self.n_plants), generator=self._gen, dtype=torch.float64)
        return torch.cat([z, u], dim=1)

    def reset(self, mask: Optional[torch.Tensor] = None) -> torch.Tensor:
        obs = self.sim.reset(mask)
        if mask is None:
            self.episode_steps.zero_()
        else:
            self.episode_steps[torch.as_tensor(mask, dtype=torch.bool, device=self.sim.device)] = 0
        return obs

    def step(self, action_idx, load_demand=None, cooling_water_temp=None, *, magnitude=None) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, Dict]:
        """action_idx: int tensor [N] of ControlAction values.  Positional order as in the reference's
        NuclearPlantEnv.step(action_idx, load_demand, cooling_water_temp) (sim.py:925-928); `load_demand` is accepted
        and, as in the reference, overridden by the plant's own power level inside the step (sim.py:158-161);
        `cooling_water_temp` (scalar or [N]) is applied before the step (sim.py:162-163); `magnitude` is keyword-only.
        Returns (observation [N,22], reward [N], done [N], info)."""
        a = torch.as_tensor(action_idx, dtype=torch.int8)
        if a.numel() != self.n_plants or int(a.min()) < 0 or int(a.max()) >= N_ACTIONS:
            raise ValueError("action_idx must hold one ControlAction value (0..14) per plant")
        if cooling_water_temp is not None:
            self.sim.state["sim.cooling_water_temp"] = cooling_water_temp
        k = self.frame_skip
        out = self.sim.step(actions=a.reshape(1, -1).expand(k, -1),
                            magnitudes=None if magnitude is None else torch.as_tensor(magnitude, dtype=torch.float64).reshape(1, -1).expand(k, -1),
                            noise=self._noise(), K=k)
        self.episode_steps += k
        done = out["done"].clone()
        obs, reward = out["observation"].clone(), out["reward"].clone()
        if k > 1:      # rewards of the substeps up to and including the first done one
            dk = out["done_k"].to(torch.float64)
            alive = (torch.cumsum(dk, dim=0) - dk) == 0
            reward = (out["reward_k"] * alive).sum(dim=0)
        info = {"episode_steps": self.episode_steps.clone(), "time_minutes": self.sim.state["sim.time_minutes"].clone()}
        if self.auto_reset and bool(done.any()):
            info["terminal_observation"] = obs[done].clone()
            obs = obs.clone()
            obs[done] = self.reset(done)[done]
        return obs, reward, done, info


def save_checkpoint(sim, path: str, maintenance=None, env=None) -> None:
    """What a resumed run needs to continue bit for bit: the SoA slab and its reset image, the parameter block, the
    step counter, threshold cooldown stamps, the in-launch monitor's stamps, the position of the device-side noise
    stream, the trajectory ring (write index and rows); with `maintenance` the work-order books of a
    BatchedAutoMaintenance (gate time, pending and executed orders, dedupe stamps); with `env` the environment's noise
    generator state and episode counters."""
    import pickle
    thr, mon, lg = sim._thr, sim._mon, sim._logged
    ck = {"slab": sim.slab.cpu(), "initial": sim._initial.cpu(), "params": torch.from_numpy(sim.params.copy()),
          "n_plants": sim.n_plants, "n_launches": sim.n_launches, "step_index": sim.step_index,
          "last_fired": None if thr is None else thr["last"].cpu(),
          "device_rng": None if getattr(sim, "_rng", None) is None else [str(v) for v in sim._rng],
          "monitor": None if mon is None else {
              "watch": list(mon["watch"]), "cap": mon["cap"], "per_substep": mon["per_substep"], "max_k": mon["max_k"],
              "watch_step": mon["watch_step"].cpu(), "first_scram": mon["first_scram"].cpu(),
              "first_nan_reset": mon["first_nan_reset"].cpu(), "status": mon["status"].cpu()},
          "ring": None if lg is None else {"names": list(lg["names"]), "rows": lg["rows"], "n": lg["n"], "ring": lg["ring"].cpu()},
          "maintenance": None, "env": None}
    if maintenance is not None:   # plain-data books; pickled bytes travel as a uint8 tensor so weights_only loading still works
        ck["maintenance"] = torch.frombuffer(bytearray(pickle.dumps(maintenance.state_dict())), dtype=torch.uint8).clone()
    if env is not None:
        ck["env"] = {"gen": env._gen.get_state(), "episode_steps": env.episode_steps.cpu()}
    torch.save(ck, path)


def load_checkpoint(path: str, device: str = "cuda:0", maintenance_table=None, maintenance=None, env=None):
    """Rebuild the simulator of save_checkpoint.  Pass the same ThresholdTable (or a fresh BatchedAutoMaintenance built on
    the returned simulator via `maintenance=lambda sim: ...`) when the checkpoint holds threshold stamps — restoring
    them without the table they index would silently change which thresholds are in cooldown, so that is an error."""
    import pickle
    from .batched import BatchedNuclearPlantSimulator
    ck = torch.load(path, map_location="cpu", weights_only=True)
    sim = BatchedNuclearPlantSimulator(int(ck["n_plants"]), ck["slab"].t().contiguous().numpy(), ck["params"].numpy(), device=device)
    sim._initial.copy_(ck["initial"].to(sim.device))
    sim.n_launches = int(ck["n_launches"])
    sim.step_index = int(ck.get("step_index", 0))
    maint = None
    if maintenance is not None:
        maint = maintenance(sim)          # BatchedAutoMaintenance.__init__ arms the thresholds on the new simulator
    elif maintenance_table is not None:
        sim.set_thresholds(maintenance_table.device_rows())
    if ck["last_fired"] is not None:
        if sim._thr is None:
            raise ValueError("checkpoint holds threshold cooldown stamps: pass maintenance_table= or maintenance=")
        sim._thr["last"].copy_(ck["last_fired"].to(sim.device))
    if ck.get("device_rng") is not None:
        seed, offset, step = (int(v) for v in ck["device_rng"])
        sim.set_device_rng(seed, plant_offset=offset, first_step=step)
    m = ck.get("monitor")
    if m is not None:
        sim.enable_monitor(watch=m["watch"], event_capacity=m["cap"], per_substep=m["per_substep"], max_k=m["max_k"])
        for k in ("watch_step", "first_scram", "first_nan_reset", "status"):
            sim._mon[k].copy_(m[k].to(sim.device))
    r = ck.get("ring")
    if r is not None:
        sim.set_logged_fields(r["names"], r["rows"])
        sim._logged["n"] = int(r["n"])
        sim._logged["ring"].copy_(r["ring"].to(sim.device))
    if ck.get("maintenance") is not None:
        if maint is None:
            raise ValueError("checkpoint holds work-order books: pass maintenance=lambda sim: BatchedAutoMaintenance(sim, table)")
        maint.load_state_dict(pickle.loads(bytes(ck["maintenance"].numpy().tobytes())))
    if env is not None and ck.get("env") is not None:
        env.sim = sim
        env._gen.set_state(ck["env"]["gen"])
        env.episode_steps = ck["env"]["episode_steps"].to(sim.device)
    return (sim, maint) if maintenance is not None else sim
