"""Public names of nuclear_sim_b200 (import-light: CUDA is touched only when a simulator is built)."""
from ._layout import N_PARAMS, N_STATE, field_index, field_names, struct_range  # noqa: F401

__all__ = ["N_STATE", "N_PARAMS", "field_names", "field_index", "struct_range",
           "BatchedNuclearPlantSimulator", "load_snapshot"]


def __getattr__(name):
    if name == "BatchedNuclearPlantSimulator":
        from .batched import BatchedNuclearPlantSimulator
        return BatchedNuclearPlantSimulator
    if name == "load_snapshot":
        from .snapshots import load_snapshot
        return load_snapshot
    raise AttributeError(name)
