"""CPU tests of the host logic of pv_plan.cu: the batched simplifySolution() passes (planning.py:195-196 -> OMPL
simplifyMax) and path.interpolate (planning.py:198).  The simplifier is the product's C++ code, reached through the test
hook pv_simplify_path_cb with the CPU ORACLE supplied as the motion validator -- on the GPU the same code asks the edge
kernel instead (tests/test_gpu_planner.py).  No device is touched here."""
import ctypes as C
import json
import os

import numpy as np
import pytest

from oracle import panda_oracle as po
from rbe550_final_project_b200 import _cabi
from rbe550_final_project_b200 import panda_model as pm
from rbe550_final_project_b200 import scenes as sc
from rbe550_final_project_b200.pathutil import interpolate, path_length

GOALS = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "goal_configs.json")))


def _wall_scene():
    wall = sc.make_obb((0.55, 0.0, 0.35), (0.5, 0.04, 0.7))
    return sc.SceneSnapshot(obb=np.array([wall], dtype=np.float32), names=["wall"], entity_idx=[1])


def _simplify(path, validator, seed=1, capacity=256):
    lib = _cabi.load()
    calls = {"batches": 0, "motions": 0}

    def cb(user, a, b, n, ok):
        qa = np.ctypeslib.as_array(a, shape=(n, 9)).astype(np.float64)
        qb = np.ctypeslib.as_array(b, shape=(n, 9)).astype(np.float64)
        res = validator(qa, qb)
        for k in range(n):
            ok[k] = 1 if res[k] else 0
        calls["batches"] += 1
        calls["motions"] += n
        return 0

    pts = np.ascontiguousarray(path, dtype=np.float64)
    out = np.empty((capacity, 9))
    n = C.c_int(0)
    counters = (C.c_int * 4)()
    rc = lib.pv_simplify_path_cb(pts.ctypes.data, len(pts), seed, _cabi.EDGE_CALLBACK(cb), None, out.ctypes.data, capacity,
                                 C.byref(n), counters)
    assert rc == 0
    return out[: n.value].copy(), [int(c) for c in counters], calls


@pytest.fixture(scope="module")
def wall_problem(c64, c32):
    snap = _wall_scene()
    scene = snap.as_oracle_scene()
    start = np.array(GOALS["goal1_scattered"]["approach_c"]["q"], dtype=np.float32)  # hand above (0.45, 0.4)
    goal = start.copy()
    goal[[0, 2, 4, 6]] *= -1  # the pose mirrored in the plane y = 0: the straight joint-space motion sweeps through the wall
    assert c64.state_margin(np.stack([start, goal]).astype(np.float64), scene).min() > 0
    assert c64.edge_margin(start[None].astype(np.float64), goal[None].astype(np.float64), scene)[0] < 0
    raws = []
    for seed in range(1, 9):
        p, it, ch = c32.rrtc(start, goal, scene, seed=seed, search=0, max_iters=4000, max_nodes=2048, max_path=128,
                             shortcut_passes=0)
        if len(p) >= 4:
            raws.append(p.astype(np.float64))
    assert len(raws) >= 3, "the CPU planner must deliver multi-vertex paths to simplify"

    def validator(qa, qb):
        return c64.edge_margin(qa, qb, scene, n_steps=0) >= 0

    return scene, raws, validator


def test_simplifier_shortens_and_stays_valid(wall_problem, c64):
    scene, raws, validator = wall_problem
    shorter = 0
    for raw in raws:
        out, counters, calls = _simplify(raw, validator, seed=3)
        assert np.array_equal(out[0], raw[0]) and np.array_equal(out[-1], raw[-1])
        assert path_length(out) <= path_length(raw) + 1e-9
        shorter += path_length(out) < 0.98 * path_length(raw)
        # every motion of the result is valid in the fp64 oracle at the planner's resolution
        assert (c64.edge_margin(out[:-1], out[1:], scene, n_steps=0) >= 0).all()
        # every pass is a BATCH: far fewer validator calls than motions validated
        assert calls["batches"] == sum(counters[:3]) and calls["motions"] == counters[3]
        assert calls["batches"] <= 13 and calls["motions"] > calls["batches"]
        again, counters2, _ = _simplify(raw, validator, seed=3)
        assert np.array_equal(out, again) and counters == counters2  # deterministic per seed
    assert shorter >= len(raws) - 1


def test_simplifier_with_the_c_validator_matches_the_python_validator(wall_problem, c32):
    """bench.py's CPU arm runs the simplifier with the oracle's C callback (no Python in the loop): same answers."""
    scene, raws, _ = wall_problem
    fn, ctx, keep = c32.edge_callback(scene)
    lib = _cabi.load()
    for raw in raws[:3]:
        ref, counters, _ = _simplify(raw, lambda qa, qb: c32.edge_margin(qa, qb, scene, n_steps=0) >= 0, seed=9)
        pts = np.ascontiguousarray(raw, dtype=np.float64)
        out = np.empty((256, 9))
        n = C.c_int(0)
        cnt = (C.c_int * 4)()
        rc = lib.pv_simplify_path_cb(pts.ctypes.data, len(pts), 9, _cabi.EDGE_CALLBACK(fn.value), C.byref(ctx),
                                     out.ctypes.data, 256, C.byref(n), cnt)
        assert rc == 0 and np.array_equal(out[: n.value], ref) and [int(c) for c in cnt] == counters
    assert ctx.motions > 0 and ctx.states >= ctx.motions


def test_simplifier_smooths_corners(wall_problem):
    """With nothing in the way, a dog-leg becomes one straight motion (vertex / partial shortcuts) -- and a corner that
    must stay (validator forbids the straight motion) is cut as far as the validator allows: a shorter path all of whose
    motions the validator accepts."""
    a = np.array(pm.Q_SAFE_HOME, dtype=np.float64)
    b, c = a.copy(), a.copy()
    b[0] += 0.8
    c[0] += 0.8
    c[1] += 0.6
    dogleg = np.stack([a, b, c])
    out, counters, _ = _simplify(dogleg, lambda qa, qb: np.ones(len(qa), bool))
    assert len(out) == 2 and np.array_equal(out[0], a) and np.array_equal(out[-1], c)

    def near_corner_only(qa, qb):
        # a motion is valid when it stays within 0.25 rad of the original dog-leg
        ok = []
        for x, y in zip(qa, qb):
            worst = 0.0
            for t in np.linspace(0, 1, 9):
                p = x + t * (y - x)
                d1 = np.linalg.norm(p - (a + np.clip(np.dot(p - a, b - a) / np.dot(b - a, b - a), 0, 1) * (b - a)))
                d2 = np.linalg.norm(p - (b + np.clip(np.dot(p - b, c - b) / np.dot(c - b, c - b), 0, 1) * (c - b)))
                worst = max(worst, min(d1, d2))
            ok.append(worst < 0.25)
        return np.array(ok)

    out, counters, _ = _simplify(dogleg, near_corner_only)
    assert len(out) >= 3 and counters[0] >= 1 and counters[1] >= 1  # partial shortcuts and B-spline steps ran
    assert path_length(out) < 0.9 * path_length(dogleg)
    assert near_corner_only(out[:-1], out[1:]).all()


def test_simplifier_edge_cases():
    a = np.array(pm.Q_SAFE_HOME, dtype=np.float64)
    b = a.copy()
    b[2] += 1.0
    never = lambda qa, qb: np.zeros(len(qa), bool)  # noqa: E731
    always = lambda qa, qb: np.ones(len(qa), bool)  # noqa: E731
    # fewer than 3 vertices: returned as they are, validator never asked
    for pts in (np.zeros((0, 9)), a[None], np.stack([a, b])):
        out, counters, calls = _simplify(pts, never)
        assert np.array_equal(out, pts) and calls["batches"] == 0
    # nothing validates: the path is unchanged
    zig = np.stack([a, b, a + 0.3, b + 0.2, a - 0.1])
    out, _, calls = _simplify(zig, never)
    assert np.array_equal(out, zig) and calls["batches"] >= 1
    # repeated vertices and zero-length segments are digested
    dup = np.stack([a, a, b, b, b, a + 0.5])
    out, _, _ = _simplify(dup, always)
    assert len(out) == 2 and np.array_equal(out[0], a) and np.array_equal(out[-1], a + 0.5)
    # a long path: candidates are sampled, vertex count is bounded by the capacity given
    rng = np.random.default_rng(0)
    long = np.cumsum(rng.normal(0, 0.05, (80, 9)), axis=0) + a
    out, counters, calls = _simplify(long, always, capacity=128)
    assert len(out) <= 128 and path_length(out) < 0.5 * path_length(long)


def test_interpolate_is_the_c_routine_and_matches_the_oracle():
    rng = np.random.default_rng(5)
    for n_pts, count in ((2, 150), (3, 150), (7, 100), (5, 5), (9, 4), (2, 2), (1, 10)):
        pts = rng.uniform(pm.Q_LOWER, pm.Q_UPPER, size=(n_pts, 9))
        got = interpolate(pts, count)
        ref = po.interpolate_path(pts, count)
        assert got.shape == ref.shape and np.abs(got - ref).max() < 1e-12
        if n_pts >= 2 and count >= n_pts:
            assert len(got) == count and np.array_equal(got[0], pts[0]) and np.array_equal(got[-1], pts[-1])
