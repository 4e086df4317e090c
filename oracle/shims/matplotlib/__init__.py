"""Test-infrastructure stub: the reference imports matplotlib unconditionally
(simulator/core/sim.py:7, data_gen/runners/maintenance_scenario_runner.py:15) but never
draws when plotting is disabled.  Any attribute resolves to a no-op."""
import sys, types


class _Anything:
    def __call__(self, *a, **k):
        return _Anything()

    def __getattr__(self, name):
        return _Anything()

    def __iter__(self):
        return iter(())


def __getattr__(name):
    return _Anything()


def use(*a, **k):
    return None


for _sub in ("pyplot", "dates", "gridspec", "patches", "ticker", "figure", "axes"):
    _m = types.ModuleType(f"matplotlib.{_sub}")
    _m.__getattr__ = lambda name: _Anything()
    sys.modules[f"matplotlib.{_sub}"] = _m
    globals()[_sub] = _m
