"""Import shim: the package directory is named ``nuclear-sim_b200`` (not a valid Python
identifier), so this module points its ``__path__`` there and re-exports the public API."""
import os as _os

__path__ = [_os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "nuclear-sim_b200")]

from ._api import N_PARAMS, N_STATE, field_index, field_names, struct_range  # noqa: F401,E402
from . import _api as _api_mod  # noqa: E402


def __getattr__(name):
    return getattr(_api_mod, name)
