#!/usr/bin/env python
"""Regenerate csrc/plant/live_fields.txt (fields read before written within one step) with the host dev tool
csrc/tools/field_liveness.cpp, from the committed snapshots.  Run after changing state.h or the physics:

    python nuclear-sim_b200/csrc/tools/live_fields.py
"""
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", "..", ".."))
sys.path.insert(0, ROOT)
from nuclear_sim_b200 import field_names, load_snapshot  # noqa: E402


def main():
    tmp = tempfile.mkdtemp()
    exe = os.path.join(tmp, "field_liveness")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-I" + os.path.join(HERE, "..", "plant"),
                           os.path.join(HERE, "field_liveness.cpp"), "-o", exe])
    files = []
    for s in ("pwr3000_oil_top_off_dt5", "pwr3000_reactor_dt1", "pwr3000_steady_dt1"):
        st, pr = load_snapshot(s)
        p = os.path.join(tmp, s + ".bin")
        np.concatenate([st, pr]).tofile(p)
        files.append(p)
    mask = subprocess.check_output([exe] + files).decode().strip()
    names = field_names()
    assert len(mask) == len(names)
    out = os.path.join(HERE, "..", "plant", "live_fields.txt")
    with open(out, "w") as fh:
        fh.write("# PlantState fields that are read before they are written within one plant_step (live on entry).\n"
                 "# Produced by csrc/tools/live_fields.py; a performance hint for the step kernel's prefetch, never a\n"
                 "# correctness input.\n")
        for n, c in zip(names, mask):
            if c == "1":
                fh.write(n + "\n")
    print(f"{mask.count('1')} of {len(mask)} fields live on entry -> {os.path.normpath(out)}")


if __name__ == "__main__":
    main()
