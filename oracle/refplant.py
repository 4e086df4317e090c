"""TEST INFRASTRUCTURE — drives the *live* Python reference (/root/reference) as the oracle.

Not importable on the GPU box (the reference is absent there): it is used in this container to
(1) validate the C/CUDA restatement step by step and (2) generate the committed golden
fixtures under tests/golden/ (see oracle/make_golden.py).

What it does:
  * puts the two shims (matplotlib stub, dataclass_wizard stand-in) and the reference on sys.path,
  * builds one reference plant per call, clearing the process-global registries the reference
    leaves behind (simulator/state/auto_register.py:136-147),
  * replaces the reference's random draws by host-supplied streams so the batched engine can
    consume identical numbers:
       - ConstantHeatSource.rng.normal(0, sigma)        (constant_heat_source.py:178)
       - np.random.normal / np.random.random in the pH controller (ph_control_system.py:278-420)
  * extracts the flat PlantState / PlantParams vectors (nuclear_sim_b200/_layout.py order) from
    the reference object graph.
"""
from __future__ import annotations

import contextlib
import io
import os
import sys
from typing import Dict, Optional

import numpy as np

REF_ROOT = os.environ.get("NPS_REFERENCE_ROOT", "/root/reference")
_HERE = os.path.dirname(os.path.abspath(__file__))
_REPO = os.path.dirname(_HERE)


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "nuclear_simulator"))


def setup_paths() -> None:
    want = [os.path.join(_HERE, "shims"),
            os.path.join(REF_ROOT, "nuclear_simulator", "data_gen"),
            os.path.join(REF_ROOT, "nuclear_simulator"),
            REF_ROOT]
    for p in reversed(want):
        if p not in sys.path:
            sys.path.insert(0, p)
    if _REPO not in sys.path:
        sys.path.insert(0, _REPO)


@contextlib.contextmanager
def quiet():
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        yield buf


class _StreamRNG:
    """Stands in for ConstantHeatSource.rng: normal(loc, scale) = loc + scale * z  with z host-supplied
    (legacy RandomState.normal is exactly loc + scale * legacy_gauss())."""

    def __init__(self):
        self.z = 0.0

    def normal(self, loc=0.0, scale=1.0):
        return loc + scale * self.z


class _PHRandom:
    """Stands in for ``np.random`` inside systems/secondary/ph_control_system.py."""

    def __init__(self):
        self.z = 0.0
        self.u = [1.0, 1.0, 1.0]
        self.n_normal = 0
        self.n_uniform = 0

    def normal(self, loc=0.0, scale=1.0):
        self.n_normal += 1
        return loc + scale * self.z

    def random(self):
        v = self.u[min(self.n_uniform, 2)]
        self.n_uniform += 1
        return v


class _NPProxy:
    def __init__(self, real, rnd):
        self._real = real
        self.random = rnd

    def __getattr__(self, name):
        return getattr(self._real, name)


class ReferencePlant:
    """One live reference plant with host-controlled random streams."""

    def __init__(self, sim, heat_rng: Optional[_StreamRNG], ph_random: Optional[_PHRandom]):
        self.sim = sim
        self.heat_rng = heat_rng
        self.ph_random = ph_random

    def step(self, action=None, magnitude: float = 1.0, noise=None):
        """noise = (z_heat, z_ph, u0, u1, u2)"""
        from systems.primary import ControlAction
        if noise is None:
            noise = (0.0, 0.0, 1.0, 1.0, 1.0)
        if self.heat_rng is not None:
            self.heat_rng.z = float(noise[0])
        if self.ph_random is not None:
            import systems.secondary.ph_control_system as phmod
            phmod.np.random = self.ph_random   # the patch is module-global: re-point it at THIS plant's stream
            self.ph_random.z = float(noise[1])
            self.ph_random.u = [float(noise[2]), float(noise[3]), float(noise[4])]
            self.ph_random.n_normal = 0
            self.ph_random.n_uniform = 0
        if action is None:
            act = ControlAction.NO_ACTION
        elif isinstance(action, (int, np.integer)):
            act = ControlAction(int(action))
        else:
            act = action
        with quiet():
            return self.sim.step(action=act, magnitude=magnitude)


def _clear_registries() -> None:
    from simulator.state.state_manager import StateManager
    for attr in ("_pending_registrations", "_global_instance_counters"):
        # _global_instance_counters numbers auto-generated ids (REA-001, WAT-00n, FEE-001 ...) per PROCESS
        # (state_manager.py:620-629); clearing it makes every plant built here look like the first plant of a fresh
        # interpreter, which is how the reference is actually run (one plant per process)
        if hasattr(StateManager, attr):
            getattr(StateManager, attr).clear()
    try:
        from simulator.state import auto_register as ar
        for name in dir(ar):
            obj = getattr(ar, name)
            if name.startswith("_instance_counters") and isinstance(obj, dict):
                obj.clear()
    except Exception:
        pass
    try:
        from simulator.state.component_metadata import ComponentRegistry
        for attr in ("_components", "_registry"):
            if hasattr(ComponentRegistry, attr) and hasattr(getattr(ComponentRegistry, attr), "clear"):
                getattr(ComponentRegistry, attr).clear()
    except Exception:
        pass
    try:
        import systems.maintenance.maintenance_orchestrator as mo
        for name in ("_orchestrator_instance", "_global_orchestrator", "_orchestrator"):
            if hasattr(mo, name):
                setattr(mo, name, None)
    except Exception:
        pass


def compose_config(action: str = "oil_top_off", duration_hours: float = 1.0) -> dict:
    setup_paths()
    with quiet():
        from config_engine.composers.comprehensive_composer import ComprehensiveComposer
        return ComprehensiveComposer().compose_action_test_scenario(target_action=action,
                                                                    duration_hours=duration_hours)


def make_reference_plant(config: Optional[dict] = None, *, dt: float = 5.0, heat_source: str = "constant",
                         noise_enabled: bool = False, noise_std_percent: float = 0.1,
                         enable_secondary: bool = True, enable_state_management: bool = False,
                         power_setpoint: Optional[float] = None) -> ReferencePlant:
    """Build a reference NuclearPlantSimulator the way MaintenanceScenarioRunner._initialize_simulator
    does (data_gen/runners/maintenance_scenario_runner.py:205-240), with controlled randomness."""
    setup_paths()
    with quiet():
        from simulator.core.sim import NuclearPlantSimulator
        from systems.primary.reactor.heat_sources import ConstantHeatSource
        from systems.primary.reactor.heat_sources.reactor_heat_source import ReactorHeatSource
        import systems.secondary.ph_control_system as phmod
        _clear_registries()
        heat_rng = None
        if heat_source == "constant":
            hs = ConstantHeatSource(rated_power_mw=3000.0, noise_enabled=noise_enabled,
                                    noise_std_percent=noise_std_percent, noise_seed=42,
                                    noise_filter_time_constant=30.0)
            heat_rng = _StreamRNG()
            hs.rng = heat_rng
        else:
            hs = ReactorHeatSource(rated_power_mw=3000.0)
        ph_random = _PHRandom()
        if not isinstance(phmod.np, _NPProxy):
            phmod.np = _NPProxy(np, ph_random)
        else:
            phmod.np.random = ph_random
        if config is None and enable_secondary:
            config = compose_config()
        sim = NuclearPlantSimulator(dt=dt, heat_source=hs, enable_secondary=enable_secondary,
                                    enable_state_management=enable_state_management,
                                    secondary_config=config if enable_secondary else None)
        if heat_source == "reactor":
            from systems.primary.reactor.reactivity_model import create_equilibrium_state
            sim.primary_physics.state = create_equilibrium_state()
            sim.state = sim.primary_physics.state
        if power_setpoint is not None and heat_source == "constant":
            hs.set_power_setpoint(power_setpoint)
    if enable_secondary:   # keep the last result dict so derived outputs can be extracted
        sec = sim.secondary_physics
        orig = sec.update_system

        def _wrapped(*a, **k):
            r = orig(*a, **k)
            sec._nps_last_result = r
            return r
        sec.update_system = _wrapped
    return ReferencePlant(sim, heat_rng, ph_random)


# ------------------------------------------------------------------------------------------
# state / parameter extraction
# ------------------------------------------------------------------------------------------
def _layout():
    setup_paths()
    import nuclear_sim_b200._layout as L
    return L


def extract_state_dict(sim) -> Dict[str, float]:
    d: Dict[str, float] = {}
    pp = sim.primary_physics
    s = pp.state
    P = "pri."
    for name in ("neutron_flux", "reactivity", "fuel_temperature", "coolant_temperature", "coolant_pressure",
                 "coolant_flow_rate", "coolant_void_fraction", "steam_temperature", "steam_pressure",
                 "steam_flow_rate", "feedwater_flow_rate", "control_rod_position", "steam_valve_position",
                 "boron_concentration", "feedwater_pump_status", "feedwater_pump_speed",
                 "feedwater_system_available", "feedwater_pump_power", "feedwater_num_running_pumps",
                 "xenon_concentration", "iodine_concentration", "samarium_concentration",
                 "burnable_poison_worth", "fuel_burnup", "power_level", "scram_status"):
        d[P + name] = float(getattr(s, name))
    for i in range(6):
        d[f"{P}precursors[{i}]"] = float(s.delayed_neutron_precursors[i])
    d[P + "thermal_power_mw"] = float(pp.thermal_power_mw)
    d[P + "total_reactivity_pcm"] = float(pp.total_reactivity_pcm)
    d[P + "scram_activated"] = float(pp.scram_activated)
    hs = pp.heat_source
    d[P + "hs_setpoint_percent"] = float(getattr(hs, "power_setpoint_percent", 100.0))
    d[P + "hs_current_power_mw"] = float(getattr(hs, "current_power_mw", 3000.0))
    d[P + "hs_time"] = float(getattr(hs, "time", 0.0))
    d[P + "hs_total_energy_mwh"] = float(getattr(hs, "total_energy_mwh", 0.0))
    d[P + "hs_filtered_noise_mw"] = float(getattr(hs, "filtered_noise_mw", 0.0))
    d[P + "hs_raw_noise_mw"] = float(getattr(hs, "raw_noise_mw", 0.0))

    S = "sim."
    if getattr(sim, "state_manager", None) is not None:
        d[S + "time_minutes"] = sim.state_manager.get_elapsed_time().total_seconds() / 60.0
    else:
        d[S + "time_minutes"] = float(getattr(sim, "time", 0.0))
    d[S + "load_demand"] = float(sim.load_demand)
    d[S + "cooling_water_temp"] = float(sim.cooling_water_temp)
    d[S + "has_last_heat_removal_factor"] = float(hasattr(sim, "_last_heat_removal_factor"))
    d[S + "last_heat_removal_factor"] = float(getattr(sim, "_last_heat_removal_factor", 0.0))
    d[S + "last_load_factor"] = float(getattr(sim, "_last_load_factor", 0.0))
    d[S + "last_feedwater_flow_factor"] = float(getattr(sim, "_last_feedwater_flow_factor", 0.0))
    d[S + "last_pump_reliability_factor"] = float(getattr(sim, "_last_pump_reliability_factor", 0.0))
    from oracle import refextract_secondary
    refextract_secondary.extract(sim, d)
    return d


def extract_params_dict(sim) -> Dict[str, float]:
    d: Dict[str, float] = {}
    hs = sim.primary_physics.heat_source
    d["dt"] = float(sim.dt)
    d["heat_source_type"] = 0.0 if type(hs).__name__ == "ConstantHeatSource" else 1.0
    d["rated_power_mw"] = float(hs.rated_power_mw)
    d["noise_enabled"] = float(getattr(hs, "noise_enabled", False))
    d["noise_std_percent"] = float(getattr(hs, "noise_std_percent", 0.0))
    d["noise_filter_time_constant"] = float(getattr(hs, "noise_filter_time_constant", 30.0))
    d["enable_secondary"] = float(bool(sim.enable_secondary and sim.secondary_physics is not None))
    from oracle import refextract_secondary
    refextract_secondary.extract_params(sim, d)
    return d


def _to_vector(d: Dict[str, float], struct: str, strict: bool = True) -> np.ndarray:
    L = _layout()
    names = L.field_names(struct)
    missing = [n for n in names if n not in d]
    if missing and strict:
        raise KeyError(f"extractor does not provide {len(missing)} {struct} fields, e.g. {missing[:8]}")
    extra = [k for k in d if k not in L.field_index(struct)]
    if extra and strict:
        raise KeyError(f"extractor provides unknown {struct} fields, e.g. {extra[:8]}")
    return np.array([d.get(n, 0.0) for n in names], dtype=np.float64)


def extract_state(sim, strict: bool = True) -> np.ndarray:
    return _to_vector(extract_state_dict(sim), "PlantState", strict)


def extract_params(sim, strict: bool = True) -> np.ndarray:
    return _to_vector(extract_params_dict(sim), "PlantParams", strict)
