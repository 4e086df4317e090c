"""BatchedNuclearPlantSimulator — N independent plants advanced in lockstep on one B200.

Keeps the step()/state/action semantics of the reference ``NuclearPlantSimulator``
(nuclear_simulator/simulator/core/sim.py:27-258) with a leading plant axis:

    sim = BatchedNuclearPlantSimulator(n_plants, initial_state, params, device="cuda:0")
    out = sim.step(actions, magnitudes, noise=noise)         # dict of tensors
    sim.state.power_level                                     # Tensor[N] view into the SoA slab

PyTorch is used for device memory, streams and (in bench.py) torch.distributed only; all physics
runs in the hand-written CUDA step kernel behind the C ABI (include/nps_b200.h).
"""
from __future__ import annotations

import ctypes
from typing import Dict, Optional, Sequence

import numpy as np
import torch

from . import _clib
from ._layout import N_PARAMS, N_STATE, field_index
from ._layout import field_names as _layout_field_names

OBS_DIM = 22
NOISE_PER_STEP = 5
NO_ACTION = 8   # ControlAction.NO_ACTION (systems/primary/__init__.py:37)


def _ptr(t: Optional[torch.Tensor]):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


class _StateView:
    """``sim.state.<field>`` -> Tensor[N] view (ReactorState attribute names map to the ``pri.`` block)."""

    def __init__(self, sim: "BatchedNuclearPlantSimulator"):
        object.__setattr__(self, "_sim", sim)

    def _resolve(self, name: str) -> int:
        ix = field_index()
        for cand in (name, "pri." + name, "sec." + name, "sim." + name):
            if cand in ix:
                return ix[cand]
        raise AttributeError(name)

    def __getattr__(self, name: str) -> torch.Tensor:
        return self._sim.slab[self._resolve(name)]

    def __setattr__(self, name: str, value) -> None:
        self._sim.slab[self._resolve(name)] = torch.as_tensor(value, dtype=torch.float64, device=self._sim.device)

    def __getitem__(self, name: str) -> torch.Tensor:
        return self._sim.slab[field_index()[name]]

    def __setitem__(self, name: str, value) -> None:
        """``sim.state["sim.cooling_water_temp"] = x`` with a scalar or an [N] array / tensor."""
        self._sim.slab[field_index()[name]] = torch.as_tensor(value, dtype=torch.float64, device=self._sim.device)


class _HeatSourceView:
    """``sim.primary_physics.heat_source.set_power_setpoint`` (constant_heat_source.py:94-102), batched."""

    def __init__(self, sim):
        self._sim = sim

    def set_power_setpoint(self, power_percent) -> None:
        sim = self._sim
        sp = torch.clamp(torch.as_tensor(power_percent, dtype=torch.float64, device=sim.device).expand(sim.n_plants), 0.0, 150.0)
        ix = field_index()
        sim.slab[ix["pri.hs_setpoint_percent"]] = sp
        sim.slab[ix["pri.hs_current_power_mw"]] = (sp / 100.0) * float(sim.params[field_index("PlantParams")["rated_power_mw"]])


class _Monitor(ctypes.Structure):
    """struct nps_monitor of include/nps_b200.h"""
    _fields_ = [("d_last_fired", ctypes.c_void_p), ("d_events", ctypes.c_void_p), ("d_n_events", ctypes.c_void_p),
                ("event_capacity", ctypes.c_uint32), ("skip_last_check", ctypes.c_int32),
                ("d_watch_fields", ctypes.c_void_p), ("d_watch_step", ctypes.c_void_p),
                ("n_watch", ctypes.c_int32), ("reserved", ctypes.c_int32),
                ("d_first_scram_step", ctypes.c_void_p), ("d_first_nan_reset_step", ctypes.c_void_p),
                ("d_status", ctypes.c_void_p), ("d_reward_k", ctypes.c_void_p), ("d_done_k", ctypes.c_void_p),
                ("step0", ctypes.c_int64)]


EVENT_DTYPE = np.dtype([("plant", np.int32), ("row", np.int32), ("step", np.int32), ("reserved", np.int32),
                        ("value", np.float64), ("time_minutes", np.float64)])   # struct nps_event

# flag fields whose first non-zero step is stamped by default (SURVEY 8 a-events e3-e12: latched trips, SG shutdown /
# replacement flags)
DEFAULT_WATCH = ("fw.pump[0].trip_active", "fw.pump[1].trip_active", "fw.pump[2].trip_active", "fw.pump[3].trip_active",
                 "fw.prot_system_trip_active", "fw.prot_npsh_low_low_trip_active", "fw.prot_npsh_critical_trip_active",
                 "turb.prot_trip_active", "cond.vs_trip_high_pressure",
                 "sgs.sg[0].tsp_shutdown_required", "sgs.sg[1].tsp_shutdown_required", "sgs.sg[2].tsp_shutdown_required",
                 "sgs.sg[0].tsp_replacement_recommended", "sgs.sg[1].tsp_replacement_recommended",
                 "sgs.sg[2].tsp_replacement_recommended", "sgs.sg[0].tif_replacement_recommended",
                 "sgs.sg[1].tif_replacement_recommended", "sgs.sg[2].tif_replacement_recommended")


class BatchedNuclearPlantSimulator:
    def __init__(self, n_plants: int, initial_state: np.ndarray, params: np.ndarray, device: str = "cuda:0"):
        if not torch.cuda.is_available():
            raise _clib.NpsError("BatchedNuclearPlantSimulator needs a CUDA device (no CPU fallback)")
        self.device = torch.device(device)
        self.n_plants = int(n_plants)
        self.L = _clib.lib()
        if self.L.nps_n_state() != N_STATE or self.L.nps_n_params() != N_PARAMS:
            raise _clib.NpsError("libnps_b200.so was built from a different state.h; rebuild")
        initial_state = np.asarray(initial_state, dtype=np.float64)
        if initial_state.ndim == 1:
            initial_state = np.broadcast_to(initial_state[None, :], (self.n_plants, N_STATE))
        if initial_state.shape != (self.n_plants, N_STATE):
            raise ValueError(f"initial_state must be [{N_STATE}] or [{self.n_plants}, {N_STATE}]")
        self.params = np.ascontiguousarray(np.asarray(params, dtype=np.float64))
        if self.params.shape != (N_PARAMS,):
            raise ValueError(f"params must be [{N_PARAMS}]")
        self.dt = float(self.params[field_index("PlantParams")["dt"]])
        dev_index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        with torch.cuda.device(dev_index):
            # SoA slab: field-major [n_state, n_plants]
            self.slab = torch.from_numpy(np.array(initial_state.T, dtype=np.float64, order="C", copy=True)).to(self.device)
            self._initial = self.slab.clone()
            h = ctypes.c_void_p()
            _clib.check(self.L.nps_create(self.n_plants, dev_index, ctypes.byref(h)))
            self._h = h
            _clib.check(self.L.nps_set_params(self._h, self.params.ctypes.data_as(ctypes.c_void_p), N_PARAMS))
            self._obs = torch.empty((OBS_DIM, self.n_plants), dtype=torch.float64, device=self.device)
            self._reward = torch.empty(self.n_plants, dtype=torch.float64, device=self.device)
            self._done = torch.empty(self.n_plants, dtype=torch.uint8, device=self.device)
        self.state = _StateView(self)
        self.heat_source = _HeatSourceView(self)
        self.n_launches = 0
        self._thr = None
        self._logged = None
        self._mon = None
        self.step_index = 0      # steps taken since construction / reset(); the in-launch monitor stamps events with it

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                self.L.nps_destroy(h)
            except Exception:
                pass
            self._h = None

    # -- stepping -----------------------------------------------------------------------------
    def _prep(self, x, dtype, k, inner=None):
        if x is None:
            return None
        t = torch.as_tensor(x, dtype=dtype, device=self.device)
        shape = (k, self.n_plants) if inner is None else (k, inner, self.n_plants)
        if t.dim() == len(shape) - 1:
            t = t.unsqueeze(0)
        if tuple(t.shape) != shape:
            t = t.expand(shape)
        return t.contiguous()

    def step(self, actions=None, magnitudes=None, noise=None, power_setpoint=None, K: int = 1,
             skip_last_check: bool = False) -> Dict[str, torch.Tensor]:
        """K fused calls of NuclearPlantSimulator.step (sim.py:130-258) for every plant.

        actions [K,N] or [N] int8 (ControlAction values; None = NO_ACTION); magnitudes [K,N] f64;
        noise [K,5,N] f64 = (z_heat, z_ph, u0, u1, u2) host-supplied streams; power_setpoint [K,N] (NaN = keep).
        Returns observation [N,22] and reward [N] of the last substep and done [N] (bool: a scram was activated in any
        substep).  With enable_monitor() every substep is observed the way the reference observes every step: event
        steps are exact whatever K is (first_scram_step, watch_steps(), drain_step_events()).
        """
        a = self._prep(actions, torch.int8, K)
        m = self._prep(magnitudes, torch.float64, K)
        z = self._prep(noise, torch.float64, K, NOISE_PER_STEP)
        sp = self._prep(power_setpoint, torch.float64, K)
        mon = self._monitor_struct(int(K), skip_last_check)
        with torch.cuda.device(self.device):
            stream = torch.cuda.current_stream(self.device).cuda_stream
            _clib.check(self.L.nps_step_monitored(self._h, _ptr(self.slab), _ptr(a), _ptr(m), _ptr(z), _ptr(sp), int(K),
                                                  _ptr(self._obs), _ptr(self._reward), _ptr(self._done),
                                                  ctypes.byref(mon) if mon is not None else None, ctypes.c_void_p(stream)))
        self.n_launches += 1
        self.step_index += int(K)
        if z is None and getattr(self, "_rng", None) is not None:
            self._rng[2] += int(K)
        out = {"observation": self._obs.t(), "reward": self._reward, "done": self._done.bool()}
        if mon is not None and self._mon["per_substep"]:
            out["reward_k"] = self._mon["reward_k"][:K]
            out["done_k"] = self._mon["done_k"][:K].bool()
        return out

    # -- in-launch monitoring (nps_step_monitored) ------------------------------------------------------------------
    def enable_monitor(self, watch: Optional[Sequence[str]] = DEFAULT_WATCH, event_capacity: Optional[int] = None,
                       per_substep: bool = False, max_k: int = 128) -> None:
        """Evaluate after EVERY substep of a fused launch what the reference evaluates after every step: first scram /
        NaN-reset step and sticky status bits per plant, the first step each `watch` flag field became non-zero, the
        threshold rows of set_thresholds() (events with their step, drained by drain_step_events()), and with
        per_substep=True the reward and done of every substep (step() then also returns reward_k / done_k [K, N])."""
        ix = field_index()
        watch = list(watch or [])
        if len(watch) > 32:
            raise ValueError("at most 32 watched fields")
        n = self.n_plants
        dev = self.device
        cap = int(event_capacity) if event_capacity else max(1 << 16, 4 * n)
        self._mon = {
            "watch": watch, "cap": cap, "per_substep": bool(per_substep), "max_k": int(max_k),
            "watch_fields": torch.as_tensor([ix[w] for w in watch] or [0], dtype=torch.int32, device=dev),
            "watch_step": torch.full((max(1, len(watch)), n), -1, dtype=torch.int32, device=dev),
            "first_scram": torch.full((n,), -1, dtype=torch.int32, device=dev),
            "first_nan_reset": torch.full((n,), -1, dtype=torch.int32, device=dev),
            "status": torch.zeros((n,), dtype=torch.int32, device=dev),
            "events": torch.zeros((cap, EVENT_DTYPE.itemsize), dtype=torch.uint8, device=dev),
            "n_events": torch.zeros((1,), dtype=torch.int32, device=dev),
            "reward_k": torch.empty((max_k, n), dtype=torch.float64, device=dev) if per_substep else None,
            "done_k": torch.empty((max_k, n), dtype=torch.uint8, device=dev) if per_substep else None,
        }

    def disable_monitor(self) -> None:
        self._mon = None

    def _monitor_struct(self, K: int, skip_last_check: bool):
        g = self._mon
        if g is None:
            return None
        if g["per_substep"] and K > g["max_k"]:
            raise ValueError(f"K={K} exceeds the monitor's max_k={g['max_k']}")
        m = _Monitor()
        thr = self._thr
        if thr is not None:
            m.d_last_fired = thr["last"].data_ptr()
            m.d_events = g["events"].data_ptr()
            m.d_n_events = g["n_events"].data_ptr()
            m.event_capacity = g["cap"]
        m.skip_last_check = 1 if skip_last_check else 0
        m.n_watch = len(g["watch"])
        m.d_watch_fields = g["watch_fields"].data_ptr()
        m.d_watch_step = g["watch_step"].data_ptr()
        m.d_first_scram_step = g["first_scram"].data_ptr()
        m.d_first_nan_reset_step = g["first_nan_reset"].data_ptr()
        m.d_status = g["status"].data_ptr()
        if g["per_substep"]:
            m.d_reward_k = g["reward_k"].data_ptr()
            m.d_done_k = g["done_k"].data_ptr()
        m.step0 = self.step_index
        return m

    def drain_step_events(self) -> np.ndarray:
        """Threshold violations recorded inside the launches since the last drain, as a structured array
        (plant, row, step, value, time_minutes) sorted by (step, plant, row); clears the device list."""
        g = self._mon
        if g is None:
            raise _clib.NpsError("enable_monitor() first")
        n = int(g["n_events"].item())
        if n == 0:
            return np.zeros(0, dtype=EVENT_DTYPE)
        if n > g["cap"]:
            raise _clib.NpsError(f"event list overflow: {n} violations, capacity {g['cap']} (enable_monitor(event_capacity=...))")
        if g.get("host") is None or g["host"].shape[0] < n:      # pinned landing buffer, grown geometrically
            g["host"] = torch.empty((max(n, 2 * (g["host"].shape[0] if g.get("host") is not None else 4096)), EVENT_DTYPE.itemsize),
                                    dtype=torch.uint8).pin_memory()
        g["host"][:n].copy_(g["events"][:n], non_blocking=True)
        g["n_events"].zero_()
        torch.cuda.current_stream(self.device).synchronize()
        ev = g["host"][:n].numpy().view(EVENT_DTYPE).reshape(-1).copy()
        key = (ev["step"].astype(np.int64) << 44) | (ev["plant"].astype(np.int64) << 12) | ev["row"].astype(np.int64)
        return ev[np.argsort(key, kind="stable")]

    @property
    def first_scram_step(self) -> torch.Tensor:
        return self._mon["first_scram"]

    @property
    def first_nan_reset_step(self) -> torch.Tensor:
        return self._mon["first_nan_reset"]

    @property
    def status(self) -> torch.Tensor:
        """Sticky per-plant status word: bit 0 a NaN reset happened (thermal_hydraulics.py:257-269), bit 1 scram latched."""
        return self._mon["status"]

    def watch_steps(self) -> Dict[str, torch.Tensor]:
        """{watched field: Tensor[N] int32 first step at which it was non-zero, -1 = never}."""
        g = self._mon
        return {w: g["watch_step"][i] for i, w in enumerate(g["watch"])}

    def set_small_batch_shape(self, shape: int) -> None:
        """Launch shape up to 18 944 plants: 0 two threads per plant (source / sink halves pipelined by one substep,
        default), 1 one thread per plant.  Bit-identical results (tests/test_gpu_parity.py)."""
        _clib.check(self.L.nps_set_small_batch_shape(self._h, int(shape)))

    def set_device_rng(self, seed: Optional[int], plant_offset: int = 0, first_step: int = 0) -> None:
        """Device-side noise (nps_set_device_rng): with a seed, step(noise=None) draws every plant-step's five random
        numbers on the GPU (Philox4x32-10, counter = (plant_offset + plant, step)); None switches it off again.
        Give each rank its first global plant id as plant_offset and the result does not depend on the sharding."""
        _clib.check(self.L.nps_set_device_rng(self._h, 0 if seed is None else 1, int(seed or 0) & (2 ** 64 - 1),
                                              int(plant_offset), int(first_step)))
        # mirrored on the host so a checkpoint can resume the stream (env.save_checkpoint): [seed, offset, next step]
        self._rng = None if seed is None else [int(seed) & (2 ** 64 - 1), int(plant_offset), int(first_step)]

    def step_host(self, h_actions, h_magnitudes, h_noise, h_setpoint, K, h_obs, h_reward, h_done) -> None:
        """Reference-facing call with HOST (pinned) per-step buffers: copies in, K substeps, copies out."""
        def hp(t):
            return ctypes.c_void_p(t.data_ptr()) if t is not None else None
        stream = torch.cuda.current_stream(self.device).cuda_stream
        _clib.check(self.L.nps_step_host(self._h, _ptr(self.slab), hp(h_actions), hp(h_magnitudes), hp(h_noise),
                                         hp(h_setpoint), int(K), hp(h_obs), hp(h_reward), hp(h_done),
                                         ctypes.c_void_p(stream)))
        self.n_launches += 1
        if h_noise is None and getattr(self, "_rng", None) is not None:
            self._rng[2] += int(K)

    def step_host_async(self, h_actions, h_magnitudes, h_noise, h_setpoint, K, h_obs, h_reward, h_done) -> int:
        """Pipelined step_host (nps_step_host_async): returns a ticket; wait(ticket) makes the outputs valid.
        Up to ``pipe_depth`` calls may be outstanding: wait for call i - pipe_depth before issuing call i."""
        def hp(t):
            return ctypes.c_void_p(t.data_ptr()) if t is not None else None
        stream = torch.cuda.current_stream(self.device).cuda_stream
        t = self.L.nps_step_host_async(self._h, _ptr(self.slab), hp(h_actions), hp(h_magnitudes), hp(h_noise), hp(h_setpoint),
                                       int(K), hp(h_obs), hp(h_reward), hp(h_done), ctypes.c_void_p(stream))
        if t < 0:
            _clib.check(t)
        self.n_launches += 1
        if h_noise is None and getattr(self, "_rng", None) is not None:
            self._rng[2] += int(K)
        return int(t)

    def wait(self, ticket: int) -> None:
        _clib.check(self.L.nps_wait(self._h, int(ticket)))

    @property
    def pipe_depth(self) -> int:
        """Launches step_host_async may have outstanding; tickets are handed out round-robin over this many slots."""
        return int(self.L.nps_pipe_depth())

    def get_observation(self) -> torch.Tensor:
        stream = torch.cuda.current_stream(self.device).cuda_stream
        _clib.check(self.L.nps_observe(self._h, _ptr(self.slab), _ptr(self._obs), _ptr(self._reward), ctypes.c_void_p(stream)))
        self.n_launches += 1
        return self._obs.t()

    def reset(self, mask: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Back to the initial state (all plants, or the masked ones).  A reset plant starts a new episode: its clock
        restarts, so its threshold cooldown stamps and monitor stamps are forgotten as well."""
        if mask is None:
            self.slab.copy_(self._initial)
            m = slice(None)
        else:
            m = torch.as_tensor(mask, dtype=torch.bool, device=self.device)
            self.slab[:, m] = self._initial[:, m]
        self._forget_episode(m)
        return self.get_observation()

    def _forget_episode(self, m) -> None:
        if self._thr is not None:
            self._thr["last"][:, m] = -float("inf")
        if self._mon is not None:
            g = self._mon
            g["watch_step"][:, m] = -1
            g["first_scram"][m] = -1
            g["first_nan_reset"][m] = -1
            g["status"][m] = 0

    # -- single-plant conveniences used by the scalar NuclearPlantSimulator facade (plant_simulator.py) ----------
    def step_plant(self, plant: int, action: int, magnitude: float, z: np.ndarray):
        """One reference-style step of a ONE-plant engine: returns (observation[22], reward, done)."""
        if self.n_plants != 1 or plant != 0:
            raise _clib.NpsError("step_plant drives a 1-plant engine; batches step in lockstep through step()")
        out = self.step(actions=torch.tensor([[int(action)]], dtype=torch.int8),
                        magnitudes=torch.tensor([[float(magnitude)]], dtype=torch.float64),
                        noise=torch.as_tensor(np.asarray(z, dtype=np.float64).reshape(1, NOISE_PER_STEP, 1)), K=1)
        return (out["observation"][0].cpu().numpy().copy(), float(out["reward"][0].item()), bool(out["done"][0].item()))

    def observe_plant(self, plant: int) -> np.ndarray:
        return self.get_observation()[plant].cpu().numpy().copy()

    def reward_plant(self, plant: int) -> float:
        """calculate_reward() of one plant's current state (nps_observe fills obs and reward together)."""
        self.get_observation()
        return float(self._reward[plant].item())

    def reset_plant(self, plant: int) -> None:
        self.slab[:, plant] = self._initial[:, plant]
        self._forget_episode(int(plant))

    def write_fields(self, plant: int, values: Dict[int, float]) -> None:
        for f, v in values.items():
            self.slab[int(f), int(plant)] = float(v)

    def synchronize(self) -> None:
        torch.cuda.synchronize(self.device)

    # -- state access -------------------------------------------------------------------------
    def current_time_minutes(self) -> float:
        """The batch clock: plants stepped in lockstep share it; after a partial reset() the reset plants' own clocks start
        again, so the largest sim.time_minutes of the batch is the one that keeps running."""
        return float(self.slab[field_index()["sim.time_minutes"]].max().item())

    def state_numpy(self) -> np.ndarray:
        """[n_plants, n_state] host copy in PlantState field order."""
        return self.slab.t().contiguous().cpu().numpy()

    def read_fields(self, names: Sequence[str]) -> np.ndarray:
        ix = field_index()
        ids = np.array([ix[n] for n in names], dtype=np.int32)
        out = np.empty((len(ids), self.n_plants), dtype=np.float64)
        stream = torch.cuda.current_stream(self.device).cuda_stream   # ordered after the launches queued on this stream
        _clib.check(self.L.nps_read_fields(self._h, _ptr(self.slab), ids.ctypes.data_as(ctypes.c_void_p), len(ids),
                                           out.ctypes.data_as(ctypes.c_void_p), ctypes.c_void_p(stream)))
        return out

    # -- threshold monitoring (state_manager.py:1307-1369) --------------------------------------
    def set_thresholds(self, rows) -> None:
        """rows: iterable of (field_name_or_None, comparator, value, cooldown_hours)."""
        cmp_code = {">": 0, "greater_than": 0, "<": 1, "less_than": 1, ">=": 2, "greater_equal": 2,
                    "<=": 3, "less_equal": 3, "==": 4, "equals": 4, "!=": 5, "not_equals": 5}
        ix = field_index()
        rows = list(rows)
        # field: PlantState field name, None (inert row) or a raw int code (>= 0 field index, <= -2 derived column)
        f = np.array([(-1 if r[0] is None else (int(r[0]) if isinstance(r[0], (int, np.integer)) else ix[r[0]])) for r in rows],
                     dtype=np.int32)
        c = np.array([cmp_code.get(r[1], 6) for r in rows], dtype=np.int32)   # unknown comparison never fires (state_manager.py:1441)
        v = np.array([r[2] for r in rows], dtype=np.float64)
        cd = np.array([r[3] * 60.0 for r in rows], dtype=np.float64)
        _clib.check(self.L.nps_set_thresholds(self._h, f.ctypes.data_as(ctypes.c_void_p), c.ctypes.data_as(ctypes.c_void_p),
                                              v.ctypes.data_as(ctypes.c_void_p), cd.ctypes.data_as(ctypes.c_void_p), len(rows)))
        n = len(rows)
        self._thr = {
            "n": n,
            "last": torch.full((n, self.n_plants), -float("inf"), dtype=torch.float64, device=self.device),
            "flags": torch.zeros(((n + 31) // 32, self.n_plants), dtype=torch.int32, device=self.device),
            "any": torch.zeros(((self.n_plants + 31) // 32,), dtype=torch.int32, device=self.device),
        }

    def clear_thresholds(self) -> None:
        empty = np.zeros(0, dtype=np.int32)
        _clib.check(self.L.nps_set_thresholds(self._h, empty.ctypes.data_as(ctypes.c_void_p), empty.ctypes.data_as(ctypes.c_void_p),
                                              empty.ctypes.data_as(ctypes.c_void_p), empty.ctypes.data_as(ctypes.c_void_p), 0))
        self._thr = None

    def check_thresholds(self):
        """Returns (flags [n_words, N] int32 bitmask tensor, any_warp [ceil(N/32)] int32 ballot words)."""
        t = self._thr
        if t is None:
            raise _clib.NpsError("set_thresholds() first")
        stream = torch.cuda.current_stream(self.device).cuda_stream
        _clib.check(self.L.nps_check_thresholds(self._h, _ptr(self.slab), _ptr(t["last"]), _ptr(t["flags"]), _ptr(t["any"]),
                                                ctypes.c_void_p(stream)))
        self.n_launches += 1
        return t["flags"], t["any"]

    def check_thresholds_events(self) -> None:
        """The threshold check of the CURRENT state with the violations appended to the monitor's event list (stamped
        with the last completed step), instead of the flag matrix: drain_step_events() returns them."""
        t, g = self._thr, self._mon
        if t is None or g is None:
            raise _clib.NpsError("set_thresholds() and enable_monitor() first")
        stream = torch.cuda.current_stream(self.device).cuda_stream
        _clib.check(self.L.nps_check_thresholds_events(self._h, _ptr(self.slab), _ptr(t["last"]), _ptr(g["events"]),
                                                       _ptr(g["n_events"]), g["cap"], self.step_index - 1, ctypes.c_void_p(stream)))
        self.n_launches += 1

    def drain_events(self):
        """Host drain: sorted list of (plant, threshold_index) that fired in the last check_thresholds().  The warp
        ballot words say which 32-plant groups have anything at all; only those columns of the flag matrix travel."""
        t = self._thr
        anyw = t["any"].cpu().numpy().view(np.uint32)
        groups = np.nonzero(anyw)[0]
        if groups.size == 0:
            return []
        lanes = (anyw[groups, None] >> np.arange(32, dtype=np.uint32)[None, :]) & 1
        plants = (groups[:, None] * 32 + np.arange(32)[None, :])[lanes.astype(bool)]
        idx = torch.as_tensor(plants, dtype=torch.long, device=self.device)
        words = t["flags"][:, idx].cpu().numpy().view(np.uint32)            # [n_words, n_flagged_plants]
        bits = (words[:, :, None] >> np.arange(32, dtype=np.uint32)[None, None, :]) & 1
        w, j, b = np.nonzero(bits)
        ev = sorted(zip(plants[j].tolist(), (w * 32 + b).tolist()))
        return ev

    def reset_cooldowns(self, plant: int, rows: Sequence[int]) -> None:
        """Forget when these thresholds of one plant last fired
        (StateManager._reset_threshold_cooldowns_for_maintenance: state_manager.py:1783-1830)."""
        self._thr["last"][torch.as_tensor(list(rows), dtype=torch.long, device=self.device), int(plant)] = -float("inf")

    def read_threshold_values(self, plants: Sequence[int], table, events=None) -> Dict:
        """{(plant, threshold index): value} — the value the violation record carries (state_manager.py:1343-1351).
        With `events` (the drained (plant, row) pairs) only those values are gathered on the device."""
        plants = list(plants)
        if not plants:
            return {}
        ix = field_index()
        if events is None:
            events = [(p, t) for p in plants for t, r in enumerate(table.rows) if r.field or r.derived]
        ev_p = torch.as_tensor([e[0] for e in events], dtype=torch.long, device=self.device)
        out = {}
        direct = [i for i, (p, t) in enumerate(events) if table.rows[t].field]
        if direct:
            f = torch.as_tensor([ix[table.rows[events[i][1]].field] for i in direct], dtype=torch.long, device=self.device)
            vals = self.slab[f, ev_p[direct]].cpu().numpy()
            for i, v in zip(direct, vals):
                out[events[i]] = float(v)
        for i, (p, t) in enumerate(events):
            r = table.rows[t]
            if r.derived == "pump_sum_wear":
                f = torch.as_tensor([ix[f"{r.unit}lub.component_wear[{c}]"] for c in range(5)], dtype=torch.long, device=self.device)
                w = self.slab[f, p].cpu().numpy()
                out[(p, t)] = float(w[0] + max(w[1], w[2], w[3]) + w[4])
        return out

    # -- maintenance effects (auto_maintenance.py:504-673 -> csrc/plant/maintenance.h) ----------------------------
    def apply_maintenance(self, requests) -> list:
        """requests: iterable of (plant, target code, action code, arg); returns the per-request status codes
        (0 failed, 1 success, 2 unsupported target), applied in order."""
        req = (np.ascontiguousarray(requests, dtype=np.int32) if isinstance(requests, np.ndarray)
               else np.array(list(requests), dtype=np.int32)).reshape(-1, 4)
        if len(req) == 0:
            return []
        cols = [np.ascontiguousarray(req[:, j]) for j in range(4)]
        status = np.zeros(len(req), dtype=np.int32)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        _clib.check(self.L.nps_apply_maintenance(self._h, _ptr(self.slab), *(c.ctypes.data_as(ctypes.c_void_p) for c in cols),
                                                 len(req), status.ctypes.data_as(ctypes.c_void_p), ctypes.c_void_p(stream)))
        self.n_launches += 1
        return status if isinstance(requests, np.ndarray) else status.tolist()

    # -- trajectory ring buffer (state_manager.py:152-233) --------------------------------------
    def set_logged_fields(self, names: Sequence[str], ring_rows: int) -> None:
        ix = field_index()
        ids = np.array([ix[n] for n in names], dtype=np.int32)
        _clib.check(self.L.nps_set_logged_fields(self._h, ids.ctypes.data_as(ctypes.c_void_p), len(ids)))
        self._logged = {"names": list(names), "rows": int(ring_rows), "n": 0,
                        "ring": torch.empty((ring_rows, len(ids), self.n_plants), dtype=torch.float64, device=self.device)}

    def set_logged_columns(self, prefix: Optional[str] = None, ring_rows: int = 256) -> int:
        """Log what the reference's CSV columns starting with `prefix` need (None: every column, e.g. "secondary.feedwater_FWP-1.");
        returns the number of fields per ring row."""
        from .export import ColumnSchema
        schema = ColumnSchema()
        which = schema.select(prefix)
        ids = schema.logged_fields(which)
        names = _layout_field_names()
        self.set_logged_fields([names[i] for i in ids], ring_rows)
        self._logged["schema"], self._logged["which"], self._logged["ids"] = schema, which, ids
        return len(ids)

    def export_trajectory(self, plant: int, filename: str, start_datetime=None, dt_minutes: Optional[float] = None) -> int:
        """The logged rows of one plant as the reference's wide CSV (same header names, same formatting as the scalar
        facade's export_to_csv); rows written."""
        import datetime as _dt
        from .export import export_ring_to_csv
        g = self._logged
        if "schema" not in g:
            raise _clib.NpsError("set_logged_columns() first")
        ring = self.drain_log()
        first = max(0, g["n"] - g["rows"]) + 1
        dt = float(self.params[field_index("PlantParams")["dt"]]) if dt_minutes is None else float(dt_minutes)
        return export_ring_to_csv(filename, ring, g["ids"], int(plant), start_datetime or _dt.datetime(2024, 1, 1), dt,
                                  first_row_step=first, schema=g["schema"], which=g["which"])

    def log_row(self) -> None:
        g = self._logged
        stream = torch.cuda.current_stream(self.device).cuda_stream
        _clib.check(self.L.nps_log_row(self._h, _ptr(self.slab), _ptr(g["ring"]), g["rows"], g["n"], ctypes.c_void_p(stream)))
        g["n"] += 1
        self.n_launches += 1

    def drain_log(self) -> np.ndarray:
        """[rows_available, n_logged, n_plants] oldest-first host copy of the ring buffer."""
        g = self._logged
        n, rows = g["n"], g["rows"]
        ring = g["ring"].cpu().numpy()
        if n <= rows:
            return ring[:n]
        start = n % rows
        return np.concatenate([ring[start:], ring[:start]], axis=0)
