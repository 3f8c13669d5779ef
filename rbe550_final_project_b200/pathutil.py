"""Host-side path post-processing that needs no validity checks (planning.py:198)."""
from __future__ import annotations

import ctypes as C
from typing import Sequence

import numpy as np

from . import _cabi


def path_length(states: Sequence[np.ndarray]) -> float:
    return float(sum(np.linalg.norm(np.asarray(b, dtype=np.float64) - np.asarray(a, dtype=np.float64))
                     for a, b in zip(states[:-1], states[1:])))


def interpolate(states: Sequence[np.ndarray], count: int) -> np.ndarray:
    """`path.interpolate(num_waypoints)` of planning.py:198, i.e. OMPL PathGeometric::interpolate(count) for
    a RealVectorStateSpace: spread `count` states over the segments in proportion to their L2 length, keeping
    every original vertex; first state = start, last = goal.  A path that already has more than `count`
    states (or fewer than 2) is returned unchanged.  One implementation: the host routine pv_plan_path itself uses
    (`pv_interpolate_path`, csrc/pv_plan.cu; plain C++ arithmetic, needs no device)."""
    pts = np.ascontiguousarray(np.asarray(states, dtype=np.float64).reshape(-1, 9))
    n_in = pts.shape[0]
    cap = max(int(count), n_in, 1)
    out = np.empty((cap, 9), dtype=np.float64)
    n = C.c_int(0)
    rc = _cabi.load().pv_interpolate_path(pts.ctypes.data, n_in, int(count), out.ctypes.data, cap, C.byref(n))
    if rc != 0:
        raise RuntimeError(f"pv_interpolate_path failed ({rc})")
    return out[: n.value].copy()
