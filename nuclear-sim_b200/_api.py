"""Public names of nuclear_sim_b200 (kept import-light: nothing here touches CUDA at import)."""
from ._layout import N_PARAMS, N_STATE, field_index, field_names, struct_range  # noqa: F401

__all__ = ["N_STATE", "N_PARAMS", "field_names", "field_index", "struct_range"]
