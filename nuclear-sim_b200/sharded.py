"""ShardedBatchedSimulator — one plant batch over the ranks of a torch.distributed job (one process per GPU).

Plants are independent, so the plant axis is cut into contiguous ranges: rank r of G owns the global
plant ids [r * N / G, (r + 1) * N / G).  Nothing is communicated on the step path; the only collective is the
all-gather of per-plant summaries at the end of a run (NCCL over NVLink on the GPUs, gloo in the CPU tests).
Everything that depends on a plant's identity is a pure function of its GLOBAL id — initial conditions
(scenarios.randomized_states), host-supplied inputs and noise (scenarios.load_following_inputs / noise_inputs), the
device-side noise stream (plant_offset of nps_set_device_rng) — so a plant's trajectory does not depend on G.
"""
from __future__ import annotations

from typing import Callable, Optional, Sequence

import numpy as np
import torch

from . import scenarios as sc


def shard_range(total_plants: int, rank: int, world: int):
    """[lo, hi) of the contiguous plant range of `rank`; the first total % world ranks own one plant more."""
    base, extra = divmod(int(total_plants), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


class ShardedBatchedSimulator:
    def __init__(self, total_plants: int, base_state: np.ndarray, params: np.ndarray, rank: Optional[int] = None,
                 world: Optional[int] = None, device: Optional[str] = None, ic_factor: float = 0.1,
                 states: Optional[Callable[[np.ndarray], np.ndarray]] = None, engine_factory=None):
        """base_state: one PlantState vector; each plant starts from scenarios.randomized_states(base_state, global id)
        (or from `states(global_ids) -> [n, n_state]` when given).  engine_factory(states, params, device) builds the
        per-rank engine (default: the CUDA BatchedNuclearPlantSimulator; the CPU tests pass the oracle stand-in)."""
        import torch.distributed as dist
        self._dist = dist if (dist.is_available() and dist.is_initialized()) else None
        self.rank = int(rank if rank is not None else (self._dist.get_rank() if self._dist else 0))
        self.world = int(world if world is not None else (self._dist.get_world_size() if self._dist else 1))
        self.total_plants = int(total_plants)
        self.lo, self.hi = shard_range(self.total_plants, self.rank, self.world)
        self.plant_ids = np.arange(self.lo, self.hi, dtype=np.int64)
        self.n_plants = len(self.plant_ids)
        st = states(self.plant_ids) if states is not None else sc.randomized_states(base_state, self.plant_ids, ic_factor)
        if engine_factory is None:
            from .batched import BatchedNuclearPlantSimulator
            dev = device or f"cuda:{torch.cuda.current_device()}"
            self.sim = BatchedNuclearPlantSimulator(self.n_plants, st, params, device=dev)
        else:
            self.sim = engine_factory(st, params, device)
        self.params = params
        self.steps_taken = 0

    # -- inputs that are functions of the global plant id -------------------------------------------------------------
    def load_following_inputs(self, k: int, t0: Optional[int] = None):
        return sc.load_following_inputs(self.plant_ids, self.steps_taken if t0 is None else t0, k)

    def noise_inputs(self, k: int, t0: Optional[int] = None, seed: int = 1000) -> np.ndarray:
        return sc.noise_inputs(self.plant_ids, self.steps_taken if t0 is None else t0, k, seed)

    def set_device_rng(self, seed: Optional[int], first_step: int = 0) -> None:
        """Device-side noise keyed by (seed, GLOBAL plant id, step): the shard's first global id is the plant offset."""
        self.sim.set_device_rng(seed, plant_offset=self.lo, first_step=first_step)

    def step(self, *args, K: int = 1, **kw):
        out = self.sim.step(*args, K=K, **kw)
        self.steps_taken += int(K)
        return out

    # -- the only collective ------------------------------------------------------------------------------------------
    def gather_summaries(self, fields: Sequence[str]) -> torch.Tensor:
        """[total_plants, len(fields)] on every rank, rows in global plant order: all_gather of this shard's
        [n_plants, len(fields)] block (shards may differ by one plant: blocks are padded to the largest shard)."""
        local = self._local_fields(fields)
        if self._dist is None or self.world == 1:
            return local
        sizes = [shard_range(self.total_plants, r, self.world) for r in range(self.world)]
        width = max(hi - lo for lo, hi in sizes)
        block = torch.zeros((width, local.shape[1]), dtype=local.dtype, device=local.device)
        block[: local.shape[0]] = local
        parts = [torch.empty_like(block) for _ in range(self.world)]
        self._dist.all_gather(parts, block)
        return torch.cat([p[: hi - lo] for p, (lo, hi) in zip(parts, sizes)])

    def _local_fields(self, fields: Sequence[str]) -> torch.Tensor:
        sim = self.sim
        if hasattr(sim, "slab"):
            from ._layout import field_index
            ix = field_index()
            idx = torch.as_tensor([ix[f] for f in fields], dtype=torch.long, device=sim.device)
            return sim.slab[idx].t().contiguous()
        from ._layout import field_index
        ix = field_index()
        return torch.from_numpy(np.ascontiguousarray(sim.state_numpy()[:, [ix[f] for f in fields]]))

    def gather_states(self) -> torch.Tensor:
        """[total_plants, n_state] on every rank (tests / small batches only: this is the whole state)."""
        from ._layout import field_names
        return self.gather_summaries(field_names("PlantState"))
