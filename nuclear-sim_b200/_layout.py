"""Derive the flat FP64 field table of ``PlantState`` / ``PlantParams`` from csrc/plant/state.h.

The structs are written in a restricted syntax (see the header of state.h) so that this small
parser is the single source of truth for field names and slab offsets on the Python side; the
C++ side gets the same order from the struct layout itself (all members are ``double``).
"""
from __future__ import annotations

import os
import re
from functools import lru_cache
from typing import Dict, List, Tuple

_HERE = os.path.dirname(os.path.abspath(__file__))
STATE_H = os.path.join(_HERE, "csrc", "plant", "state.h")

_MEMBER = re.compile(r"^\s*(\w+)\s+(\w+)((?:\[\w+\])*)\s*;")
_STRUCT = re.compile(r"^\s*struct\s+(\w+)\s*\{")
_CONST = re.compile(r"^\s*(?:static\s+)?constexpr\s+int\s+(\w+)\s*=\s*(\d+)\s*;")


_DISCRETE: Dict[Tuple[str, str], bool] = {}    # (struct, member) carries the `// @discrete` annotation


def _parse(path: str) -> Tuple[Dict[str, list], Dict[str, int]]:
    structs: Dict[str, list] = {}
    consts: Dict[str, int] = {}
    cur = None
    with open(path) as fh:
        for raw in fh:
            line = raw.split("//", 1)[0]
            m = _CONST.match(line)
            if m:
                consts[m.group(1)] = int(m.group(2))
                continue
            m = _STRUCT.match(line)
            if m:
                cur = m.group(1)
                structs[cur] = []
                continue
            if cur is None:
                continue
            if line.strip().startswith("};"):
                cur = None
                continue
            m = _MEMBER.match(line)
            if m:
                typ, name, dims = m.groups()
                shape = [consts[d] if d in consts else int(d) for d in re.findall(r"\[(\w+)\]", dims)]
                structs[cur].append((typ, name, shape))
                _DISCRETE[(cur, name)] = "@discrete" in raw.split("//", 1)[1] if "//" in raw else False
    return structs, consts


def _flatten(structs, typ: str, prefix: str, out: List[str]) -> None:
    for mtyp, name, shape in structs[typ]:
        idxs = [""]
        for d in shape:
            idxs = [f"{p}[{i}]" for p in idxs for i in range(d)]
        for ix in idxs:
            full = f"{prefix}{name}{ix}"
            if mtyp == "double":
                out.append(full)
            else:
                _flatten(structs, mtyp, full + ".", out)


@lru_cache(maxsize=None)
def field_names(struct: str = "PlantState") -> Tuple[str, ...]:
    structs, _ = _parse(STATE_H)
    out: List[str] = []
    _flatten(structs, struct, "", out)
    return tuple(out)


@lru_cache(maxsize=None)
def field_index(struct: str = "PlantState") -> Dict[str, int]:
    return {n: i for i, n in enumerate(field_names(struct))}


@lru_cache(maxsize=None)
def discrete_field_names(struct: str = "PlantState") -> Tuple[str, ...]:
    """Fields that hold flags / enums / counters / latches (exact small integers stored as doubles): the members
    annotated `// @discrete` in state.h.  The parity tests compare these bit for bit, everything else to tolerance."""
    structs, _ = _parse(STATE_H)
    out: List[str] = []

    def walk(typ: str, prefix: str) -> None:
        for mtyp, name, shape in structs[typ]:
            idxs = [""]
            for d in shape:
                idxs = [f"{p}[{i}]" for p in idxs for i in range(d)]
            for ix in idxs:
                full = f"{prefix}{name}{ix}"
                if mtyp == "double":
                    if _DISCRETE.get((typ, name), False):
                        out.append(full)
                else:
                    walk(mtyp, full + ".")
    walk(struct, "")
    return tuple(out)


def struct_range(member_path: str, struct: str = "PlantState") -> Tuple[int, int]:
    """[lo, hi) flat index range covered by a (possibly nested) member such as ``"sg[1]"``."""
    names = field_names(struct)
    pre = member_path + "."
    idx = [i for i, n in enumerate(names) if n == member_path or n.startswith(pre) or n.startswith(member_path + "[")]
    if not idx:
        raise KeyError(member_path)
    return idx[0], idx[-1] + 1


N_STATE = len(field_names("PlantState"))
N_PARAMS = len(field_names("PlantParams"))
