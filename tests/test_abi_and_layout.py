"""CPU: the C-ABI library loads and exports every symbol include/nps_b200.h declares; the Python layout
parser, the generated field table and the compiled structs agree; creating a handle without a GPU fails loudly."""
import ctypes
import os
import re

import pytest

from tests import _util as U

ROOT = U.ROOT


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "nps_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(nps_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported():
    from nuclear_sim_b200 import _clib
    L = _clib.lib()
    declared = _declared_symbols()
    assert declared, "no declarations parsed"
    for sym in declared:
        assert hasattr(L, sym), f"libnps_b200.so does not export {sym}"
    assert sorted(_clib.EXPORTED_SYMBOLS) == declared


def test_layout_agrees_with_compiled_structs(oracle_lib):
    from nuclear_sim_b200 import N_PARAMS, N_STATE, _clib, field_names
    L = _clib.lib()
    assert L.nps_n_state() == N_STATE == oracle_lib.nps_oracle_n_state()
    assert L.nps_n_params() == N_PARAMS == oracle_lib.nps_oracle_n_params()
    names = field_names("PlantState")
    assert [L.nps_field_name(i).decode() for i in range(N_STATE)] == list(names)
    pn = field_names("PlantParams")
    assert [L.nps_param_name(i).decode() for i in range(N_PARAMS)] == list(pn)
    assert L.nps_field_name(N_STATE) is None


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from nuclear_sim_b200 import _clib
    L = _clib.lib()
    h = ctypes.c_void_p()
    assert L.nps_create(4, 0, ctypes.byref(h)) != 0
    assert b"no CUDA device" in L.nps_last_error()
    from nuclear_sim_b200 import BatchedNuclearPlantSimulator, load_snapshot
    s, p = load_snapshot()
    with pytest.raises(_clib.NpsError):
        BatchedNuclearPlantSimulator(4, s, p)


def test_product_does_not_touch_oracle():
    """The product path must never import / link / call anything under oracle/."""
    pkg = os.path.join(ROOT, "nuclear-sim_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".h", ".cuh", ".inc")):
                text = open(os.path.join(dirpath, f)).read()
                assert "libnps_oracle" not in text and "cpu_port" not in text, f"{f} references the oracle library"
                for line in text.splitlines():
                    if re.match(r"\s*(from|import)\s+oracle\b", line):
                        raise AssertionError(f"{f} imports oracle: {line}")


def test_plain_c_example_compiles_and_links(tmp_path):
    """examples/step_from_c.c builds against include/nps_b200.h and links against the in-tree library with nothing but
    gcc and the CUDA runtime: the boundary really is a C ABI (no compute call here - there is no GPU)."""
    import os
    import shutil
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    lib_dir = os.path.join(root, "nuclear-sim_b200", "_lib")
    if not (shutil.which("gcc") and os.path.exists("/usr/local/cuda/include/cuda_runtime_api.h")
            and os.path.exists(os.path.join(lib_dir, "libnps_b200.so"))):
        import pytest
        pytest.skip("gcc / CUDA headers / built library not present")
    exe = str(tmp_path / "step_from_c")
    subprocess.check_call(["gcc", "-O2", "-Wall", "-Werror", "-I", os.path.join(root, "include"), "-I", "/usr/local/cuda/include",
                           os.path.join(root, "examples", "step_from_c.c"), "-o", exe, "-L", lib_dir, "-lnps_b200",
                           "-L", "/usr/local/cuda/lib64", "-lcudart", f"-Wl,-rpath,{lib_dir}"])
    assert os.path.getsize(exe) > 0


def test_discrete_fields_are_an_explicit_list_and_hold_integers():
    """The bit-exact class of fields is declared in state.h (`// @discrete`), not guessed from names: 92 flat fields;
    in every state the live reference produced for the fixtures they hold exact integers, and every field whose name
    says flag / status / trip / alarm is on the list."""
    import glob
    import numpy as np
    from nuclear_sim_b200 import field_index, field_names
    from nuclear_sim_b200._layout import discrete_field_names
    disc = discrete_field_names()
    assert len(disc) == 92 and len(set(disc)) == 92
    ix = field_index()
    cols = np.array([ix[n] for n in disc])
    n_checked = 0
    for path in sorted(glob.glob(os.path.join(U.GOLDEN, "*.npz"))):
        z = np.load(path, allow_pickle=False)
        for key in ("states", "state0", "before", "after"):
            if key in z.files:
                v = z[key].reshape(-1, z[key].shape[-1])[:, cols]
                assert np.array_equal(v, np.round(v)), f"{os.path.basename(path)}:{key} has a non-integer discrete field"
                n_checked += v.shape[0]
    assert n_checked > 1000
    for n in field_names("PlantState"):
        if any(k in n for k in ("trip_active", "scram", "_alarm", "is_operating", "shutdown_required")):
            assert n in disc, n
