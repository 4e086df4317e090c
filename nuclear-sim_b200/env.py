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
    def __init__(self, sim, auto_reset: bool = True, seed: int = 0):
        self.sim = sim
        self.n_plants = sim.n_plants
        self.action_space_size = N_ACTIONS
        self.observation_space_size = 22
        self.auto_reset = auto_reset
        self._gen = torch.Generator(device="cpu").manual_seed(seed)
        self.episode_steps = torch.zeros(self.n_plants, dtype=torch.int64, device=sim.device)

    def _noise(self) -> torch.Tensor:
        z = torch.randn((1, 2, self.n_plants), generator=self._gen, dtype=torch.float64)
        u = torch.rand((1, 3, self.n_plants), generator=self._gen, dtype=torch.float64)
        return torch.cat([z, u], dim=1)

    def reset(self, mask: Optional[torch.Tensor] = None) -> torch.Tensor:
        obs = self.sim.reset(mask)
        if mask is None:
            self.episode_steps.zero_()
        else:
            self.episode_steps[torch.as_tensor(mask, dtype=torch.bool, device=self.sim.device)] = 0
        return obs

    def step(self, action_idx, magnitude=None, cooling_water_temp=None) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, Dict]:
        """action_idx: int tensor [N] of ControlAction values.  Returns (observation [N,22], reward [N], done [N], info)."""
        a = torch.as_tensor(action_idx, dtype=torch.int8)
        if a.numel() != self.n_plants or int(a.min()) < 0 or int(a.max()) >= N_ACTIONS:
            raise ValueError("action_idx must hold one ControlAction value (0..14) per plant")
        if cooling_water_temp is not None:
            self.sim.state["sim.cooling_water_temp"] = cooling_water_temp
        out = self.sim.step(actions=a.reshape(1, -1), magnitudes=None if magnitude is None else torch.as_tensor(magnitude, dtype=torch.float64).reshape(1, -1),
                            noise=self._noise(), K=1)
        self.episode_steps += 1
        done = out["done"].clone()
        obs, reward = out["observation"].clone(), out["reward"].clone()
        info = {"episode_steps": self.episode_steps.clone(), "time_minutes": self.sim.state["sim.time_minutes"].clone()}
        if self.auto_reset and bool(done.any()):
            info["terminal_observation"] = obs[done].clone()
            obs = obs.clone()
            obs[done] = self.reset(done)[done]
        return obs, reward, done, info


def save_checkpoint(sim, path: str, maintenance=None) -> None:
    """Everything needed to resume a batch bit for bit: the SoA slab, the parameter block, threshold cooldown stamps,
    and the position of the device-side noise stream when that mode is on."""
    thr = sim._thr
    torch.save({"slab": sim.slab.cpu(), "initial": sim._initial.cpu(), "params": torch.from_numpy(sim.params.copy()),
                "n_plants": sim.n_plants, "n_launches": sim.n_launches,
                "last_fired": None if thr is None else thr["last"].cpu(),
                "device_rng": None if getattr(sim, "_rng", None) is None else [str(v) for v in sim._rng],
                "maintenance_last_check": None if maintenance is None else maintenance.last_check_time}, path)


def load_checkpoint(path: str, device: str = "cuda:0", maintenance_table=None):
    from .batched import BatchedNuclearPlantSimulator
    ck = torch.load(path, map_location="cpu", weights_only=True)
    sim = BatchedNuclearPlantSimulator(int(ck["n_plants"]), ck["slab"].t().contiguous().numpy(), ck["params"].numpy(), device=device)
    sim._initial.copy_(ck["initial"].to(sim.device))
    sim.n_launches = int(ck["n_launches"])
    if maintenance_table is not None:
        sim.set_thresholds(maintenance_table.device_rows())
    if ck["last_fired"] is not None and sim._thr is not None:
        sim._thr["last"].copy_(ck["last_fired"].to(sim.device))
    if ck.get("device_rng") is not None:
        seed, offset, step = (int(v) for v in ck["device_rng"])
        sim.set_device_rng(seed, plant_offset=offset, first_step=step)
    return sim
