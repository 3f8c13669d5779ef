"""Host-side path post-processing that needs no validity checks (planning.py:198)."""
from __future__ import annotations

import math
from typing import List, Sequence

import numpy as np


def path_length(states: Sequence[np.ndarray]) -> float:
    return float(sum(np.linalg.norm(np.asarray(b, dtype=np.float64) - np.asarray(a, dtype=np.float64))
                     for a, b in zip(states[:-1], states[1:])))


def interpolate(states: Sequence[np.ndarray], count: int) -> np.ndarray:
    """`path.interpolate(num_waypoints)` of planning.py:198, i.e. OMPL PathGeometric::interpolate(count) for
    a RealVectorStateSpace: spread `count` states over the segments in proportion to their L2 length, keeping
    every original vertex; first state = start, last = goal.  A path that already has more than `count`
    states (or fewer than 2) is returned unchanged."""
    pts = [np.asarray(s, dtype=np.float64) for s in states]
    n_in = len(pts)
    if count < n_in or n_in < 2:
        return np.array(pts)
    seg = [float(np.linalg.norm(pts[i + 1] - pts[i])) for i in range(n_in - 1)]
    remaining = float(sum(seg))
    budget = int(count)
    out: List[np.ndarray] = []
    last = n_in - 1
    for i in range(last):
        a, b = pts[i], pts[i + 1]
        out.append(a)
        room = budget + i - n_in  # interior states this segment may still take
        if room > 0:
            if i + 1 == last:
                want = room + 2
            elif remaining > 0.0:
                want = int(math.floor(0.5 + budget * seg[i] / remaining)) + 1
            else:
                want = 2
            inner = 0
            if want > 2:
                inner = min(want - 2, room)
                t = (np.arange(1, inner + 1, dtype=np.float64) / (inner + 1))[:, None]
                out.extend(a + t * (b - a))
            budget -= inner + 1
            remaining -= seg[i]
        else:
            budget -= 1
    out.append(pts[last])
    return np.array(out)
