"""Helpers shared by the tests: load tests/golden/*.npz into the packed layouts of include/vap.h."""
from __future__ import annotations

import glob
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
NA, APA = 12, 4
F_REVERSE, F_STOP, F_TANGENT = 1, 2, 4


def case_names():
    return sorted(os.path.basename(p)[5:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "case_*.npz")))


def load_case(name):
    g = dict(np.load(os.path.join(GOLDEN_DIR, f"case_{name}.npz")))
    n = len(g["points_ft"])
    na = np.zeros((n, NA))
    na[:, 0:2] = g["points_ft"]
    na[:, 2] = g["n_turn"]
    na[:, 3] = g["n_wait"]
    na[:, 4] = g["n_maxvel"]
    na[:, 5] = g["n_maxacc"]
    na[:, 6:8] = g["n_tangent"]
    na[:, 8] = g["n_inmag"]
    na[:, 9] = g["n_outmag"]
    na[:, 10] = g["rot_cos"]
    na[:, 11] = g["rot_sin"]
    nf = (g["n_reverse"].astype(np.int32) * F_REVERSE + g["n_stop"].astype(np.int32) * F_STOP
          + g["n_has_tangent"].astype(np.int32) * F_TANGENT).astype(np.int32)
    A = len(g["ap_t"])
    apa = np.zeros((A, APA))
    apf = np.zeros(A, dtype=np.int32)
    if A:
        apa[:, 0] = g["ap_t"]
        apa[:, 1] = g["ap_wait"]
        apa[:, 2] = g["ap_maxvel"]
        apa[:, 3] = g["ap_maxacc"]
        apf[:] = g["ap_stop"].astype(np.int32) * F_STOP
    g["node_attr"], g["node_flags"], g["ap_attr"], g["ap_flags"] = na, nf, apa, apf
    g["n"] = n
    return g


def ulp_diff(a, b):
    """Distance in units of the last place between two float64 arrays (same-sign finite values)."""
    a = np.ascontiguousarray(a, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    ia = a.view(np.int64).copy()
    ib = b.view(np.int64).copy()
    ia[ia < 0] = np.int64(-(2 ** 63)) - ia[ia < 0]
    ib[ib < 0] = np.int64(-(2 ** 63)) - ib[ib < 0]
    return np.abs(ia - ib)


def bit_equal(a, b):
    a = np.ascontiguousarray(a, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    return a.shape == b.shape and bool(np.all(a.view(np.int64) == b.view(np.int64)))
