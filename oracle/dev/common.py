"""Dev helpers (test infrastructure): load the host oracle lib, compare flat state vectors."""
import ctypes, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import refplant as R  # noqa: E402

LIB = ctypes.CDLL(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "_build", "libnps_oracle.so"))


def ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def dvec(x):
    return np.ascontiguousarray(np.asarray(x, dtype=np.float64))


def canonicalize(v):
    """Rotate the pH deviation ring buffer so that its oldest entry sits at index 0 (representation only)."""
    L = R._layout(); ix = L.field_index()
    v = v.copy()
    h = int(v[ix['ph.dev_head']]); b = ix['ph.dev_hist[0]']
    if h:
        v[b:b + 100] = np.roll(v[b:b + 100], -h); v[ix['ph.dev_head']] = 0.0
    return v


def compare(c, ref, prefix=None, tol=0.0, top=12, names=None):
    L = R._layout()
    c = canonicalize(c); ref = canonicalize(ref)
    names = names or L.field_names()
    err = np.abs(c - ref) / np.maximum(np.abs(ref), 1e-300)
    err[c == ref] = 0.0
    err[np.isnan(c) & np.isnan(ref)] = 0.0
    idx = [i for i, n in enumerate(names) if (prefix is None or n.startswith(prefix))]
    bad = [(err[i], names[i], c[i], ref[i]) for i in idx if not (err[i] <= tol)]
    bad.sort(key=lambda t: -t[0] if t[0] == t[0] else -1e99)
    return bad[:top], (max((err[i] for i in idx), default=0.0))
