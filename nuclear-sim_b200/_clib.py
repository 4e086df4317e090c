"""ctypes binding of the C ABI in include/nps_b200.h (libnps_b200.so, built in-tree).

There is no CPU fallback: importing this module without the built library, or creating a handle
without a CUDA device, raises.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_int, c_int64, c_void_p

from . import _build

_LIB = None


class NpsError(RuntimeError):
    pass


def lib() -> ctypes.CDLL:
    global _LIB
    if _LIB is not None:
        return _LIB
    path = os.environ.get("NPS_B200_LIB", _build.LIB_PATH)   # tuning variants only; default is the in-tree build
    if not os.path.exists(path):
        raise NpsError(f"{path} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                       "(nvcc, sm_100a). nuclear_sim_b200 has no CPU fallback.")
    L = ctypes.CDLL(path)
    L.nps_abi_version.restype = c_int
    L.nps_last_error.restype = c_char_p
    L.nps_n_state.restype = c_int
    L.nps_n_params.restype = c_int
    L.nps_field_name.restype = c_char_p
    L.nps_field_name.argtypes = [c_int]
    L.nps_param_name.restype = c_char_p
    L.nps_param_name.argtypes = [c_int]
    L.nps_create.argtypes = [c_int64, c_int, POINTER(c_void_p)]
    L.nps_destroy.argtypes = [c_void_p]
    L.nps_destroy.restype = None
    L.nps_n_plants.argtypes = [c_void_p]
    L.nps_n_plants.restype = c_int64
    L.nps_set_params.argtypes = [c_void_p, c_void_p, c_int]
    L.nps_step.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p,
                           c_void_p, c_void_p]
    L.nps_step_host.argtypes = L.nps_step.argtypes
    L.nps_step_monitored.argtypes = L.nps_step.argtypes[:-1] + [c_void_p, c_void_p]
    L.nps_step_host_async.argtypes = L.nps_step.argtypes
    L.nps_wait.argtypes = [c_void_p, c_int]
    L.nps_set_device_rng.argtypes = [c_void_p, c_int, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint64]
    L.nps_device_rng_draws.argtypes = [ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint64, c_void_p]
    L.nps_selftest_pow.argtypes = [c_void_p, c_void_p, c_void_p, c_int64, c_void_p]
    L.nps_measure_fp64_peak.argtypes = [c_int, c_int, c_void_p, c_void_p]
    L.nps_set_small_batch_shape.argtypes = [c_void_p, c_int]
    L.nps_pipe_depth.argtypes = []
    L.nps_pipe_depth.restype = c_int
    L.nps_observe.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]
    L.nps_set_thresholds.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int]
    L.nps_check_thresholds.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]
    L.nps_check_thresholds_events.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, ctypes.c_uint32, ctypes.c_int32, c_void_p]
    L.nps_set_logged_fields.argtypes = [c_void_p, c_void_p, c_int]
    L.nps_log_row.argtypes = [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_void_p]
    L.nps_read_fields.argtypes = [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p]
    L.nps_n_maintenance_actions.restype = c_int
    L.nps_maintenance_action_name.restype = c_char_p
    L.nps_maintenance_action_name.argtypes = [c_int]
    L.nps_apply_maintenance.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p]
    # host-side work-order table (csrc/nps_workorders.cpp)
    c_double = ctypes.c_double
    L.nps_wo_create.argtypes = [c_int64, c_int, c_int, c_int] + [c_void_p] * 8 + [c_double, c_int, POINTER(c_void_p)]
    L.nps_wo_destroy.argtypes = [c_void_p]
    L.nps_wo_destroy.restype = None
    L.nps_wo_group.argtypes = [c_void_p, c_int64] + [c_void_p] * 11
    L.nps_wo_group.restype = c_int64
    L.nps_wo_issue.argtypes = [c_void_p, c_double, c_int64] + [c_void_p] * 6 + [c_int64, c_void_p, c_void_p]
    L.nps_wo_issue.restype = c_int64
    L.nps_wo_n_pending.argtypes = [c_void_p]
    L.nps_wo_n_pending.restype = c_int64
    L.nps_wo_due.argtypes = [c_void_p, c_double, c_int64] + [c_void_p] * 8
    L.nps_wo_due.restype = c_int64
    L.nps_wo_complete.argtypes = [c_void_p]
    L.nps_wo_reset_plants.argtypes = [c_void_p, c_void_p, c_int64]
    L.nps_wo_sizes.argtypes = [c_void_p, c_void_p, c_void_p]
    L.nps_wo_export.argtypes = [c_void_p] + [c_void_p] * 5
    L.nps_wo_import.argtypes = [c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p]
    if L.nps_abi_version() != 2:
        raise NpsError("libnps_b200.so ABI version mismatch")
    _LIB = L
    return L


def check(rc: int) -> None:
    if rc != 0:
        raise NpsError(lib().nps_last_error().decode())


EXPORTED_SYMBOLS = [
    "nps_abi_version", "nps_last_error", "nps_n_state", "nps_n_params", "nps_field_name", "nps_param_name",
    "nps_create", "nps_destroy", "nps_n_plants", "nps_set_params", "nps_step", "nps_step_monitored", "nps_step_host", "nps_step_host_async",
    "nps_wait", "nps_pipe_depth", "nps_set_small_batch_shape", "nps_measure_fp64_peak", "nps_set_device_rng", "nps_device_rng_draws", "nps_selftest_pow", "nps_observe",
    "nps_set_thresholds", "nps_check_thresholds", "nps_check_thresholds_events", "nps_set_logged_fields", "nps_log_row", "nps_read_fields",
    "nps_n_maintenance_actions", "nps_maintenance_action_name", "nps_apply_maintenance",
    "nps_wo_create", "nps_wo_destroy", "nps_wo_group", "nps_wo_issue", "nps_wo_n_pending", "nps_wo_due", "nps_wo_complete",
    "nps_wo_reset_plants", "nps_wo_sizes", "nps_wo_export", "nps_wo_import",
]
