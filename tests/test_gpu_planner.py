"""GPU tests of the drop-in planner: PlannerInterface.plan_path contract and the batched
RRT-Connect front end.  Paths are re-validated with the edge kernel AND the CPU oracle."""
import collections.abc
import json
import os

import numpy as np
import pytest
import torch

from conftest import random_configs
from rbe550_final_project_b200 import panda_model as pm
from rbe550_final_project_b200 import scenes as sc
from rbe550_final_project_b200.planning import PlannerInterface, PlanningError, SUPPORTED_PLANNERS, Waypoints
from rbe550_final_project_b200.validity import PandaValidityError
from rbe550_final_project_b200.sim_stub import create_scene
from rbe550_final_project_b200.validity import unpack_bits

pytestmark = pytest.mark.gpu

GOALS = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "goal_configs.json")))


def _path_ok(pv, c64, scene, path, attached=-1):
    """every state valid, every segment valid at OMPL's resolution -- on the GPU and in the oracle"""
    p = np.asarray(path, dtype=np.float32)
    w = pv.check_states_host(p)
    assert unpack_bits(w, len(p)).all()
    e = pv.check_edges_host(p[:-1], p[1:], n_steps=0)
    assert unpack_bits(e, len(p) - 1).all()
    m = c64.edge_margin(p[:-1].astype(np.float64), p[1:].astype(np.float64), scene.as_oracle_scene(), n_steps=0,
                        attached=attached)
    assert (m > -1e-4).all()


def test_plan_path_contract(pv, c64):
    scene, franka, blocks = create_scene("goal1_scattered")
    franka.set_qpos(pm.Q_SAFE_HOME)
    planner = PlannerInterface(franka, scene, validity=pv)
    calls_before = franka.raw.set_qpos_calls
    for name in ("approach_r", "grasp_r", "approach_c"):
        goal = np.array(GOALS["goal1_scattered"][name]["q"])
        path = planner.plan_path(qpos_goal=goal, num_waypoints=150, attached_object=None, timeout=10.0)
        # a sequence of 150 CPU fp32 tensors of shape (9,), rows of one (150, 9) block (planning.py:200, 232-242)
        assert isinstance(path, collections.abc.Sequence) and isinstance(path, Waypoints) and len(path) == 150 and path
        assert path.tensor.shape == (150, 9) and path[3].data_ptr() == path.tensor[3].data_ptr()
        assert all(isinstance(w, torch.Tensor) and w.device.type == "cpu" and w.dtype == torch.float32
                   and w.shape == (9,) for w in path)
        arr = np.stack([w.numpy() for w in path])
        assert np.allclose(arr[0], pm.Q_SAFE_HOME, atol=1e-6) and np.allclose(arr[-1], goal, atol=1e-6)
        # caller-side conversions of motion_primitives.py:164-176 work
        assert np.array(path[-1], dtype=float).shape == (9,)
        _path_ok(pv, c64, sc.goal1_scattered(), arr)
        st = planner.last_stats
        assert st["solved"] and st["validated"] == 1 and st["attempts"] == 1
        # these three goals are straight-line motions: the line and its 150 waypoints are validated by ONE launch
        assert st["vertices"] == 2 and st["speculative_hit"] == 1 and st["launches"] == 1
        assert planner.validate_trajectory(path).all()
        assert path[-1].tolist() == arr[-1].tolist() and [w.tolist() for w in path[:2]] == arr[:2].tolist()
    # the same seed gives the same plan, whatever the 32 racing searches do
    planner.rng_seed = 77
    a = planner.plan_path(qpos_goal=goal, num_waypoints=150)
    planner.rng_seed = 77
    b = planner.plan_path(qpos_goal=goal, num_waypoints=150)
    assert a == b and np.array_equal(a.array, b.array)
    # the robot is put back where it was (planning.py:205) and never moved in between
    assert franka.raw.set_qpos_calls == calls_before + 5
    assert np.allclose(franka.get_qpos(), pm.Q_SAFE_HOME)


def test_plan_path_soft_failures_return_empty(pv):
    scene, franka, _ = create_scene("goal1_scattered")
    franka.set_qpos(pm.Q_SAFE_HOME)
    planner = PlannerInterface(franka, scene, validity=pv)
    in_table = np.array([0, 1.7, 0, -0.1, 0, 0.5, 0, 0.04, 0.04])
    assert planner.plan_path(qpos_goal=in_table, timeout=1.0) == []
    out_of_bounds = pm.Q_SAFE_HOME.copy()
    out_of_bounds[7:] = 0.0405  # the README's finger-drift case: start above the 0.04 limit
    assert planner.plan_path(qpos_goal=pm.Q_SCENE_INIT, qpos_start=out_of_bounds, timeout=1.0) == []


def test_plan_path_hard_errors(pv):
    scene, franka, _ = create_scene("goal1_scattered")
    planner = PlannerInterface(franka, scene, validity=pv)
    with pytest.raises(PlanningError):
        planner.plan_path(qpos_goal=pm.Q_SAFE_HOME, planner="LazyPRM")  # not one of planning.py:108-117
    # names the reference accepts but the device has no kernel for: answered by RRTConnect unless strict
    franka.set_qpos(pm.Q_SCENE_INIT)
    assert len(planner.plan_path(qpos_goal=pm.Q_SAFE_HOME, planner="BITstar", num_waypoints=20)) == 20
    with pytest.raises(PlanningError):
        PlannerInterface(franka, scene, validity=pv, strict_planners=True).plan_path(qpos_goal=pm.Q_SAFE_HOME, planner="BITstar")
    with pytest.raises(PlanningError):
        planner.plan_path(qpos_goal=pm.Q_SAFE_HOME[:7])
    franka.raw._solver.n_envs = 4
    with pytest.raises(PlanningError):
        planner.plan_path(qpos_goal=pm.Q_SAFE_HOME)
    franka.raw._solver.n_envs = 0
    assert "RRTConnect" in SUPPORTED_PLANNERS


def test_attached_object_forgiveness(pv, c64):
    """planning.py:221-230: hand/finger contacts with the attached block are forgiven, nothing else."""
    scene, franka, blocks = create_scene("goal1_scattered")
    franka.set_qpos(pm.Q_SAFE_HOME)
    planner = PlannerInterface(franka, scene, validity=pv)
    planner.refresh_scene()
    # grasp pose with the fingers closed onto block r (idx 1): fingers penetrate the block
    q = np.array(GOALS["goal1_scattered"]["grasp_r"]["q"])
    q[7:] = 0.015
    pv.set_attached(-1)
    assert not pv.is_state_valid(q)
    pv.set_attached(planner._snapshot.index_of_entity(blocks["r"].idx))
    assert pv.is_state_valid(q)
    pv.set_attached(planner._snapshot.index_of_entity(blocks["g"].idx))
    assert not pv.is_state_valid(q)
    pv.set_attached(-1)
    # and through plan_path: planning away from that pose only works with the block attached
    assert planner.plan_path(qpos_goal=pm.Q_SAFE_HOME, qpos_start=q, timeout=1.0) == []
    path = planner.plan_path(qpos_goal=pm.Q_SAFE_HOME, qpos_start=q, attached_object=blocks["r"], timeout=10.0,
                             num_waypoints=50)
    assert len(path) == 50
    assert pv.attached == -1  # plan_path leaves the (shared) handle with nothing attached ...
    pv.set_attached(0)        # ... so the re-check of the path sets the grasp itself
    _path_ok(pv, c64, sc.goal1_scattered(), np.stack([w.numpy() for w in path]), attached=0)
    pv.set_attached(-1)
    # the planner object itself keeps attached_object (planning.py:153) and applies it per query
    assert planner.attached_object is blocks["r"] and planner._is_ompl_state_valid(q) is True
    planner.attached_object = None
    assert planner._is_ompl_state_valid(q) is False


def test_validity_callback_shape(pv):
    scene, franka, _ = create_scene("goal3_tower")
    planner = PlannerInterface(franka, scene, validity=pv)
    chk = planner.state_validity_checker()
    assert chk(list(pm.Q_SAFE_HOME)) is True
    assert planner._is_ompl_state_valid(pm.Q_SAFE_HOME) is True
    assert chk([0, 1.7, 0, -0.1, 0, 0.5, 0, 0.04, 0.04]) is False
    assert chk.calls == 2
    assert planner.check_motion(pm.Q_SAFE_HOME, pm.Q_SCENE_INIT) in (True, False)


def test_batched_rrtc_tower_scene(pv, c64):
    """BASELINE config 4 in miniature: many start/goal pairs in the tall-tower scene."""
    snap = sc.goal3_tower()
    pv.set_scene(snap)
    pv.set_attached(-1)
    cand = random_configs(4000, 123)
    ok = unpack_bits(pv.check_states_host(cand), len(cand))
    valid = cand[ok]
    nq = 512
    starts, goals = valid[:nq], valid[nq:2 * nq]
    paths, plen, iters, checks = pv.rrtc_batch(starts, goals, max_iters=2000, max_nodes=2048, max_path=128, seed=5,
                                               replicas=1, shortcut_passes=2)
    solved = plen > 0
    assert solved.mean() > 0.9, f"success rate {solved.mean():.3f}"
    assert (checks[solved] > 0).all() and (iters[solved] >= 1).all()
    for k in np.nonzero(solved)[0][:64]:
        p = paths[k, : plen[k]]
        assert np.array_equal(p[0], starts[k]) and np.array_equal(p[-1], goals[k])
        _path_ok(pv, c64, snap, p)
    # determinism with replicas = 1
    paths2, plen2, _, _ = pv.rrtc_batch(starts, goals, max_iters=2000, max_nodes=2048, max_path=128, seed=5,
                                        replicas=1, shortcut_passes=2)
    assert np.array_equal(plen, plen2) and np.array_equal(paths, paths2)
    # replicas: OR-parallel searches still return valid paths
    paths3, plen3, _, _ = pv.rrtc_batch(starts[:64], goals[:64], seed=9, replicas=8)
    assert (plen3 > 0).mean() >= solved[:64].mean() - 0.05
    for k in np.nonzero(plen3 > 0)[0][:16]:
        _path_ok(pv, c64, snap, paths3[k, : plen3[k]])


def test_device_rrtc_follows_the_oracle_planner(pv, c32):
    """Same seeds, same sample stream, same tie breaks: the device planner and the CPU restatement must take the same
    decisions, so paths are bit-identical unless a state inside the 1e-4 m contact band flipped a verdict."""
    snap = sc.goal3_tower()
    pv.set_scene(snap)
    pv.set_attached(-1)
    pv.set_flags(True, False)
    cand = random_configs(3000, 321)
    valid = cand[unpack_bits(pv.check_states_host(cand), len(cand))]
    nq = 160
    starts, goals = valid[:nq], valid[nq:2 * nq]
    paths, plen, iters, checks = pv.rrtc_batch(starts, goals, max_iters=500, max_nodes=1024, max_path=128, seed=11,
                                               replicas=1, shortcut_passes=2)
    same, solved_both = 0, 0
    for k in range(nq):
        p, it, ch = c32.rrtc(starts[k], goals[k], snap.as_oracle_scene(), seed=11, search=k, max_iters=500,
                             max_nodes=1024, max_path=128, shortcut_passes=2)
        if len(p) == plen[k] and np.array_equal(p, paths[k, : plen[k]]) and it == iters[k] and ch == checks[k]:
            same += 1
        if len(p) > 0 and plen[k] > 0:
            solved_both += 1
    assert same >= int(0.97 * nq), f"only {same}/{nq} searches identical to the oracle planner"
    assert solved_both >= int(0.9 * nq)


def test_rrtc_batch_is_split_invariant(pv):
    """The random stream of a search is keyed by its GLOBAL query id (query_offset + row): a batch planned in shards --
    other calls, other GPUs (distributed.rrtc_batch_sharded) -- gives exactly the rows of the unsplit call."""
    from rbe550_final_project_b200.distributed import rrtc_batch_sharded
    wall = sc.make_obb((0.55, 0.0, 0.35), (0.5, 0.04, 0.7))
    snap = sc.SceneSnapshot(obb=np.array([wall], dtype=np.float32), names=["wall"], entity_idx=[1])
    pv.set_scene(snap)
    pv.set_attached(-1)
    pv.set_flags(True, False)
    quat = np.array([0.0, 1.0, 0.0, 0.0])
    q_left, ok1, _ = pv.ik_batch(np.array([[0.5, 0.3, 0.3]]), quat[None], pm.Q_SAFE_HOME, n_seeds=128)
    q_right, ok2, _ = pv.ik_batch(np.array([[0.5, -0.3, 0.3]]), quat[None], pm.Q_SAFE_HOME, n_seeds=128)
    assert ok1[0] and ok2[0]
    # the same hard query 101 times: the rows differ only through their random streams
    nq = 101
    starts, goals = np.repeat(q_left, nq, axis=0), np.repeat(q_right, nq, axis=0)
    kw = dict(max_iters=2000, max_nodes=2048, max_path=96, seed=13, replicas=1, shortcut_passes=2)
    full = pv.rrtc_batch(starts, goals, **kw)
    assert (full[1] > 0).mean() > 0.9 and len(set(full[2].tolist())) > 3, "the searches must really sample"
    used = np.arange(full[0].shape[1])[None, :] < full[1][:, None]
    for cuts in ([0, 50, nq], [0, 1, 33, 34, 100, nq]):
        parts = [pv.rrtc_batch(starts[a:b], goals[a:b], query_offset=a, **kw) for a, b in zip(cuts[:-1], cuts[1:])]
        for j in (1, 2, 3):
            assert np.array_equal(np.concatenate([p[j] for p in parts]), full[j]), (cuts, j)
        got = np.concatenate([p[0] for p in parts])
        assert np.array_equal(got[used], full[0][used]), cuts  # rows beyond a path's length are scratch
    # without the offset a shard draws the streams of rows 0.. instead of 50..
    other = pv.rrtc_batch(starts[50:], goals[50:], **kw)
    assert np.array_equal(other[2], full[2][: nq - 50]) and not np.array_equal(other[2], full[2][50:])
    # world size 1: the sharded front end is the plain call
    one = rrtc_batch_sharded(pv, starts, goals, packed=True, **kw)
    assert np.array_equal(one[0], full[0][used])
    for j in (1, 2, 3):
        assert np.array_equal(one[j], full[j])


def test_validate_trajectory(pv, c64):
    scene, franka, blocks = create_scene("goal3_tower")
    franka.set_qpos(pm.Q_SAFE_HOME)
    planner = PlannerInterface(franka, scene, validity=pv)
    goal = np.array(GOALS["goal3_tower"]["approach_top"]["q"])
    path = planner.plan_path(qpos_goal=goal, num_waypoints=60, timeout=10.0)
    ok = planner.validate_trajectory(path)
    assert ok.shape == (59,) and ok.all()
    # an un-planned joint-space lerp straight through the tower is caught
    through = np.array(GOALS["goal3_tower"]["approach_top"]["q"])
    low = through.copy()
    low[1] += 0.6   # shoulder forward: the hand sweeps down into the tower
    lerp = [through + t * (low - through) for t in np.linspace(0, 1, 20)]
    got = planner.validate_trajectory(lerp)
    assert not got.all()
    # segment by segment against the fp64 oracle (OMPL's DiscreteMotionValidator restated), outside the contact band
    pts = np.asarray(lerp, dtype=np.float32).astype(np.float64)
    oscene = sc.goal3_tower().as_oracle_scene()
    m = c64.edge_margin(pts[:-1], pts[1:], oscene, n_steps=0)
    m[0] = min(m[0], c64.state_margin(pts[:1], oscene)[0])  # segment 0 also answers for waypoint 0
    clear = np.abs(m) > 1e-4
    assert clear.sum() >= 15 and np.array_equal(got[clear], m[clear] >= 0)


def test_nontrivial_planning_around_a_wall(pv, c64, c32):
    """A wall between start and goal forces real tree growth (several iterations, multi-segment paths): the device
    planner must still return oracle-valid paths and follow the CPU planner decision for decision."""
    wall = sc.make_obb((0.55, 0.0, 0.35), (0.5, 0.04, 0.7))
    snap = sc.SceneSnapshot(obb=np.array([wall], dtype=np.float32), names=["wall"], entity_idx=[1])
    pv.set_scene(snap)
    pv.set_attached(-1)
    pv.set_flags(True, False)
    quat = np.array([0.0, 1.0, 0.0, 0.0])
    q_left, ok1, _ = pv.ik_batch(np.array([[0.5, 0.3, 0.3]]), quat[None], pm.Q_SAFE_HOME, n_seeds=128)
    q_right, ok2, _ = pv.ik_batch(np.array([[0.5, -0.3, 0.3]]), quat[None], pm.Q_SAFE_HOME, n_seeds=128)
    assert ok1[0] and ok2[0]
    # the straight joint-space segment crosses the wall
    assert not unpack_bits(pv.check_edges_host(q_left, q_right, n_steps=0), 1)[0]
    nq = 64
    starts, goals = np.repeat(q_left, nq, axis=0), np.repeat(q_right, nq, axis=0)
    paths, plen, iters, checks = pv.rrtc_batch(starts, goals, max_iters=2000, max_nodes=2048, max_path=128, seed=21,
                                               replicas=1, shortcut_passes=2)
    assert (plen > 0).mean() > 0.95
    assert np.median(iters[plen > 0]) >= 2 and (plen[plen > 0] >= 3).all()
    same = 0
    for k in range(nq):
        if plen[k] == 0:
            continue
        p = paths[k, : plen[k]]
        _path_ok(pv, c64, snap, p)
        if k < 24:
            po_path, it, ch = c32.rrtc(starts[k], goals[k], snap.as_oracle_scene(), seed=21, search=k, max_iters=2000,
                                       max_nodes=2048, max_path=128, shortcut_passes=2)
            same += int(len(po_path) == plen[k] and np.array_equal(po_path, p) and it == iters[k])
    assert same >= 21, f"only {same}/24 searches identical to the CPU planner"
    # and through the drop-in, with OR-parallel replicas
    scene, franka, _ = create_scene("goal1_scattered")
    franka.set_qpos(q_left[0])
    planner = PlannerInterface(franka, snap, validity=pv)
    planner.rng_seed = 5
    path = planner.plan_path(qpos_goal=q_right[0], num_waypoints=200, timeout=10.0)
    st = planner.last_stats
    # the 200 resampled waypoints are checked more finely than the planner's 1 % resolution inside pv_plan_path, and a
    # path that grazes the wall between the planner's samples is replaced / re-planned: EVERY segment is valid
    assert len(path) == 200 and planner.validate_trajectory(path).all() and st["validated"] == 1
    arr = path.array.astype(np.float64)
    assert (c64.edge_margin(arr[:-1], arr[1:], snap.as_oracle_scene(), n_steps=0) > -1e-4).all()
    assert st["vertices_raw"] >= 3 and st["partial_rounds"] + st["bspline_steps"] + st["reduce_rounds"] >= 1
    # deterministic: the winner among the 32 searches is the one with the smallest (iterations, replica id)
    planner.rng_seed = 5
    again = planner.plan_path(qpos_goal=q_right[0], num_waypoints=200, timeout=10.0)
    assert np.array_equal(again.array, path.array)


def test_rrtc_capacity_limits_fail_cleanly(pv):
    """Exhausted path / node / iteration budgets give 'no solution' (length 0), never garbage."""
    wall = sc.make_obb((0.55, 0.0, 0.35), (0.5, 0.04, 0.7))
    snap = sc.SceneSnapshot(obb=np.array([wall], dtype=np.float32), names=["wall"], entity_idx=[1])
    pv.set_scene(snap)
    pv.set_attached(-1)
    quat = np.array([[0.0, 1.0, 0.0, 0.0]])
    ql, _, _ = pv.ik_batch(np.array([[0.5, 0.3, 0.3]]), quat, pm.Q_SAFE_HOME, n_seeds=100)  # rounded up to 128
    qr, _, _ = pv.ik_batch(np.array([[0.5, -0.3, 0.3]]), quat, pm.Q_SAFE_HOME, n_seeds=100)
    starts, goals = np.repeat(ql, 16, axis=0), np.repeat(qr, 16, axis=0)
    for bad in (dict(max_path=1), dict(max_nodes=7), dict(max_nodes=-5), dict(replicas=257), dict(max_iters=-1)):
        args = dict(max_iters=2000, max_nodes=2048, max_path=128, seed=3, replicas=1, shortcut_passes=2)
        args.update(bad)
        with pytest.raises(PandaValidityError):  # out-of-range capacities are refused, never replaced by defaults
            pv.rrtc_batch(starts, goals, **args)
    for kw in (dict(max_path=2), dict(max_nodes=8), dict(max_iters=1)):
        args = dict(max_iters=2000, max_nodes=2048, max_path=128, seed=3, replicas=1, shortcut_passes=2)
        args.update(kw)
        paths, plen, iters, checks = pv.rrtc_batch(starts, goals, **args)
        if "max_nodes" in kw:
            # tiny trees may still connect; whatever comes back must be a valid path, and some searches must give up
            assert (plen == 0).any()
            for k in np.nonzero(plen > 0)[0]:
                p = paths[k, : plen[k]]
                assert np.array_equal(p[0], starts[k]) and np.array_equal(p[-1], goals[k])
                assert unpack_bits(pv.check_edges_host(p[:-1], p[1:], n_steps=0), len(p) - 1).all()
        else:
            assert (plen == 0).all(), kw
    # invalid end points are reported through the iteration count
    bad = np.array([[0, 1.7, 0, -0.1, 0, 0.5, 0, 0.04, 0.04]], dtype=np.float32)
    oob = ql.copy()
    oob[0, 7] = 0.05
    _, plen, iters, _ = pv.rrtc_batch(np.concatenate([bad, ql, oob]), np.concatenate([qr, bad, qr]), check_endpoints=True)
    assert list(plen) == [0, 0, 0] and list(iters) == [-1, -2, -1]


def test_unknown_attached_entity_forgives_nothing(pv):
    scene, franka, blocks = create_scene("goal1_scattered")
    planner = PlannerInterface(franka, scene, validity=pv)
    planner.refresh_scene()

    class Ghost:
        idx = 99

    assert planner._attached_index(Ghost()) == -1
    assert planner._attached_index(None) == -1
    assert planner._attached_index(blocks["m"]) == 4


def test_simplify_solution_shortens_and_stays_valid(pv, c64):
    """ss.simplifySolution() (planning.py:196): the batched simplifyMax passes of pv_plan.cu, validated by the edge kernel."""
    from rbe550_final_project_b200.pathutil import path_length
    wall = sc.make_obb((0.55, 0.0, 0.35), (0.5, 0.04, 0.7))
    snap = sc.SceneSnapshot(obb=np.array([wall], dtype=np.float32), names=["wall"], entity_idx=[1])
    pv.set_scene(snap)
    pv.set_attached(-1)
    quat = np.array([[0.0, 1.0, 0.0, 0.0]])
    ql, _, _ = pv.ik_batch(np.array([[0.5, 0.3, 0.3]]), quat, pm.Q_SAFE_HOME, n_seeds=128)
    qr, _, _ = pv.ik_batch(np.array([[0.5, -0.3, 0.3]]), quat, pm.Q_SAFE_HOME, n_seeds=128)
    scene, franka, _ = create_scene("goal1_scattered")
    franka.set_qpos(ql[0])
    planner = PlannerInterface(franka, snap, validity=pv)
    planner.refresh_scene()
    shorter = 0
    for seed in range(12):
        paths, plen, _, _ = pv.rrtc_batch(ql, qr, seed=40 + seed, replicas=1, shortcut_passes=0)
        assert plen[0] >= 3
        raw = paths[0, : plen[0]].astype(np.float64)
        cut = planner.simplify_path(raw, seed=seed)
        assert np.array_equal(cut[0], raw[0]) and np.array_equal(cut[-1], raw[-1])
        assert path_length(cut) <= path_length(raw) + 1e-9
        shorter += path_length(cut) < path_length(raw) - 1e-6
        _path_ok(pv, c64, snap, cut)
        assert np.array_equal(cut, planner.simplify_path(raw, seed=seed))
        assert sum(pv.last_simplify_counters[k] for k in ("partial_rounds", "bspline_steps", "reduce_rounds")) >= 1
    assert shorter >= 10
    # smooth_path=False returns the raw tree path resampled; both variants are fully valid trajectories
    for smooth in (True, False):
        path = planner.plan_path(qpos_goal=qr[0], num_waypoints=120, smooth_path=smooth, timeout=10.0)
        assert len(path) >= 120 and planner.validate_trajectory(path).all()
        assert (planner.last_stats["partial_rounds"] + planner.last_stats["bspline_steps"] > 0) == smooth or \
            planner.last_stats["vertices_raw"] == 2


def test_validate_trajectory_checks_the_first_waypoint(pv):
    """a motion check assumes its start valid, so waypoint 0 must be checked as a state."""
    scene, franka, _ = create_scene("goal1_scattered")
    planner = PlannerInterface(franka, scene, validity=pv)
    bad = np.array([0, 1.7, 0, -0.1, 0, 0.5, 0, 0.04, 0.04])  # folded into the table
    ok = planner.validate_trajectory([bad, pm.Q_SAFE_HOME, pm.Q_SCENE_INIT])
    assert ok.tolist() == [False, True]
    ok = planner.validate_trajectory([pm.Q_SAFE_HOME, pm.Q_SCENE_INIT, pm.Q_SAFE_HOME])
    assert ok.tolist() == [True, True]


def test_replicas_are_deterministic_and_split_invariant(pv):
    """The winner of a query is the search with the smallest (iterations, replica id): the same on every run, and the
    same whether the batch is planned in one call or in shards (VERDICT r1 weak 11 / ADVICE)."""
    wall = sc.make_obb((0.55, 0.0, 0.35), (0.5, 0.04, 0.7))
    snap = sc.SceneSnapshot(obb=np.array([wall], dtype=np.float32), names=["wall"], entity_idx=[1])
    pv.set_scene(snap)
    pv.set_attached(-1)
    quat = np.array([[0.0, 1.0, 0.0, 0.0]])
    ql, _, _ = pv.ik_batch(np.array([[0.5, 0.3, 0.3]]), quat, pm.Q_SAFE_HOME, n_seeds=128)
    qr, _, _ = pv.ik_batch(np.array([[0.5, -0.3, 0.3]]), quat, pm.Q_SAFE_HOME, n_seeds=128)
    nq = 48
    starts, goals = np.repeat(ql, nq, axis=0), np.repeat(qr, nq, axis=0)
    kw = dict(max_iters=2000, max_nodes=2048, max_path=128, seed=17, replicas=8, shortcut_passes=2)
    runs = [pv.rrtc_batch(starts, goals, **kw) for _ in range(4)]
    used = np.arange(128)[None, :] < runs[0][1][:, None]
    assert (runs[0][1] > 0).all() and len(set(runs[0][2].tolist())) > 3
    for r in runs[1:]:
        assert all(np.array_equal(r[j], runs[0][j]) for j in (1, 2, 3)) and np.array_equal(r[0][used], runs[0][0][used])
    parts = [pv.rrtc_batch(starts[a:b], goals[a:b], query_offset=a, **kw) for a, b in ((0, 7), (7, 8), (8, nq))]
    for j in (1, 2, 3):
        assert np.array_equal(np.concatenate([p[j] for p in parts]), runs[0][j])
    assert np.array_equal(np.concatenate([p[0] for p in parts])[used], runs[0][0][used])
    # more replicas never need MORE iterations than the best of fewer: replica 0 of 8 is the single search
    single = pv.rrtc_batch(starts, goals, **dict(kw, replicas=1))
    keys8 = runs[0][2]
    # (the single search of query k has global id k; with 8 replicas the ids are 8k..8k+7, so only the winner's
    # optimality can be compared: its iteration count is minimal over ITS replicas, not over other seeds)
    assert np.median(keys8) <= np.median(single[2])


def test_one_million_queries_bounded_memory(pv):
    """1 048 576 queries in one call -- the device arena is bounded by chunking, the paths come back
    packed -- and the packed rows equal the dense form on a slice."""
    snap = sc.goal3_tower()
    pv.set_scene(snap)
    pv.set_attached(-1)
    cand = random_configs(40_000, 99)
    valid = cand[unpack_bits(pv.check_states_host(cand), len(cand))]
    rng = np.random.default_rng(0)
    nq = 1 << 20
    starts = valid[rng.integers(0, len(valid), nq)]
    goals = valid[rng.integers(0, len(valid), nq)]
    free0 = torch.cuda.mem_get_info()[0]
    states, off, plen, iters, checks = pv.rrtc_batch(starts, goals, seed=3, replicas=1, packed=True)
    used_gb = (free0 - torch.cuda.mem_get_info()[0]) / 2**30
    assert used_gb < 4.0, f"the planner's arena grew by {used_gb:.1f} GB"
    assert (plen >= 2).mean() > 0.999 and len(states) == int(plen.sum())
    assert np.array_equal(off, np.concatenate([[0], np.cumsum(plen)[:-1]]))
    solved = np.nonzero(plen >= 2)[0]
    assert np.array_equal(states[off[solved]], starts[solved]) and np.array_equal(states[off[solved] + plen[solved] - 1], goals[solved])
    sl = slice(500_000, 500_000 + 3000)
    dense = pv.rrtc_batch(starts[sl], goals[sl], seed=3, replicas=1, query_offset=sl.start, max_path=128)
    assert np.array_equal(dense[1], plen[sl]) and np.array_equal(dense[2], iters[sl]) and np.array_equal(dense[3], checks[sl])
    for k in range(0, 3000, 37):
        g = sl.start + k
        assert np.array_equal(dense[0][k, : plen[g]], states[off[g]: off[g] + plen[g]])


def test_single_tree_rrt_planner(pv, c64, c32):
    """planner="RRT" (planning.py:108-117 lists it; og.RRT defaults: 5 % goal bias, same range): device == CPU restatement,
    paths valid, and plan_path accepts the name."""
    wall = sc.make_obb((0.55, 0.0, 0.35), (0.5, 0.04, 0.7))
    snap = sc.SceneSnapshot(obb=np.array([wall], dtype=np.float32), names=["wall"], entity_idx=[1])
    pv.set_scene(snap)
    pv.set_attached(-1)
    quat = np.array([[0.0, 1.0, 0.0, 0.0]])
    ql, _, _ = pv.ik_batch(np.array([[0.5, 0.3, 0.3]]), quat, pm.Q_SAFE_HOME, n_seeds=128)
    qr, _, _ = pv.ik_batch(np.array([[0.5, -0.3, 0.3]]), quat, pm.Q_SAFE_HOME, n_seeds=128)
    nq = 32
    starts, goals = np.repeat(ql, nq, axis=0), np.repeat(qr, nq, axis=0)
    paths, plen, iters, checks = pv.rrtc_batch(starts, goals, max_iters=4000, max_nodes=4096, max_path=128, seed=77,
                                               replicas=1, shortcut_passes=2, planner="RRT")
    assert (plen > 0).mean() > 0.7  # a single tree in 9-D needs ~1500 iterations here (median); some searches run out
    same = 0
    for k in range(nq):
        if plen[k] == 0:
            continue
        p = paths[k, : plen[k]]
        assert np.array_equal(p[0], starts[k]) and np.array_equal(p[-1], goals[k])
        _path_ok(pv, c64, snap, p)
        if k < 12:
            po_path, it, ch = c32.rrtc(starts[k], goals[k], snap.as_oracle_scene(), seed=77, search=k, max_iters=4000,
                                       max_nodes=4096, max_path=128, shortcut_passes=2, planner="RRT")
            same += int(len(po_path) == plen[k] and np.array_equal(po_path, p) and it == iters[k])
    assert same >= 10
    # the single tree needs more iterations than the bidirectional search on the same problem
    _, plen_c, iters_c, _ = pv.rrtc_batch(starts, goals, max_iters=4000, max_nodes=4096, max_path=128, seed=77, replicas=1)
    assert np.median(iters[plen > 0]) >= np.median(iters_c[plen_c > 0])
    scene, franka, _ = create_scene("goal1_scattered")
    franka.set_qpos(ql[0])
    planner = PlannerInterface(franka, snap, validity=pv)
    path = planner.plan_path(qpos_goal=qr[0], num_waypoints=100, planner="RRT", timeout=10.0)
    # the resampled waypoints fall between the states the planner's validator sampled (1 % resolution, App. D), so a
    # path that grazes the wall edge can show a few waypoint-level contacts: the finer check must agree almost everywhere
    assert len(path) >= 100 and planner.validate_trajectory(path).all()
    planner.strict_planners = True
    with pytest.raises(PlanningError):
        planner.plan_path(qpos_goal=qr[0], planner="PRM")


def test_config4_full_size(pv, c64):
    """BASELINE config 4 at its full size: 4096 start/goal pairs (valid, hand above 0.15 m) in the tall-tower scene, one
    launch.  Every query is solved, every path starts and ends where asked, and a sample of the paths is valid in the fp64
    oracle at the planner's own resolution."""
    snap = sc.goal3_tower()
    pv.set_scene(snap)
    pv.set_attached(-1)
    pv.set_flags(True, False)
    cand = random_configs(60000, 4096)
    ok = unpack_bits(pv.check_states_host(cand), len(cand))
    ok &= pv.fk(torch.as_tensor(cand, device="cuda")).cpu().numpy()[:, 8, 2] > 0.15
    valid = cand[ok]
    nq = 4096
    assert len(valid) >= 2 * nq
    starts, goals = valid[:nq], valid[nq:2 * nq]
    paths, plen, iters, checks = pv.rrtc_batch(starts, goals, max_iters=2000, max_nodes=2048, max_path=128, seed=7,
                                               replicas=1, shortcut_passes=2)
    assert (plen >= 2).all(), f"{(plen < 2).sum()} of {nq} queries unsolved"
    idx = np.arange(nq)
    assert np.array_equal(paths[idx, 0], starts) and np.array_equal(paths[idx, plen - 1], goals)
    assert (iters >= 1).all() and (checks > 0).all()
    rng = np.random.default_rng(1)
    hard = np.argsort(-iters)[:40]  # the searches that had to grow trees, plus a random sample
    for k in np.concatenate([hard, rng.choice(nq, 60, replace=False)]):
        _path_ok(pv, c64, snap, paths[k, : plen[k]])


def test_nn_candidates_of_a_sharded_tree(pv):
    """pv_nn_candidates + pv_rrtc_steer: a tree dealt node by node to `world` ranks (emulated on one GPU), the per-rank
    candidates stacked the way the all-gather delivers them: the reduction picks the brute-force nearest node (ties to
    the lowest global index), and the motion is the nearest node steered towards the target by at most `range`."""
    dev = pv.device
    rng = np.random.default_rng(77)
    T, cap_global, n_pairs = 5, 300, 64
    sizes_g = np.array([1, 2, 37, 300, 129])
    nodes = rng.uniform(pm.Q_LOWER, pm.Q_UPPER, size=(T, cap_global, 9)).astype(np.float32)
    nodes[3, 17] = nodes[3, 250]  # an exact tie between two ranks' nodes: the lower global index must win
    tree_of = rng.integers(0, T, size=n_pairs).astype(np.int32)
    targets = rng.uniform(pm.Q_LOWER, pm.Q_UPPER, size=(n_pairs, 9)).astype(np.float32)
    targets[0] = nodes[3, 250]
    tree_of[0] = 3
    for world in (1, 2, 3, 8):
        slots = (cap_global + world - 1) // world
        cands = []
        for rank in range(world):
            loc = np.zeros((T, 9, slots), np.float32)
            lsz = np.zeros(T, np.int32)
            for t in range(T):
                mine = nodes[t, rank:sizes_g[t]:world]
                lsz[t] = len(mine)
                loc[t, :, : len(mine)] = mine.T
            cands.append(pv.nn_candidates(torch.from_numpy(loc).to(dev), torch.from_numpy(lsz).to(dev),
                                          torch.from_numpy(tree_of).to(dev), torch.from_numpy(targets).to(dev), rank, world))
        # the fused form stores the same records rank-major into a "peer" buffer (here: one local buffer as the only peer)
        gathered = torch.zeros((world, n_pairs, 11), dtype=torch.float32, device=dev)
        peer_tab = torch.tensor([gathered.data_ptr()], dtype=torch.int64, device=dev)
        for rank in range(world):
            loc = np.zeros((T, 9, slots), np.float32)
            lsz = np.zeros(T, np.int32)
            for t in range(T):
                mine = nodes[t, rank:sizes_g[t]:world]
                lsz[t] = len(mine)
                loc[t, :, : len(mine)] = mine.T
            pv.nn_candidates_gather(torch.from_numpy(loc).to(dev), torch.from_numpy(lsz).to(dev),
                                    torch.from_numpy(tree_of).to(dev), torch.from_numpy(targets).to(dev), rank, world,
                                    peer_tab.data_ptr(), 0, n_peers=1)
        torch.cuda.synchronize()
        assert torch.equal(gathered.view(torch.int32), torch.stack(cands).contiguous().view(torch.int32)), world
        gi, ea, eb, reach = pv.rrtc_steer(gathered, torch.from_numpy(targets).to(dev), 1.0)
        gi, ea, eb, reach = gi.cpu().numpy(), ea.cpu().numpy(), eb.cpu().numpy(), reach.cpu().numpy()
        for i in range(n_pairs):
            t = tree_of[i]
            d2 = ((nodes[t, : sizes_g[t]].astype(np.float64) - targets[i]) ** 2).sum(1)
            best = int(np.argmin(d2))
            assert d2[gi[i]] <= d2[best] * (1 + 1e-6) + 1e-12, (world, i)
            assert np.array_equal(ea[i], nodes[t, gi[i]])
            d = np.sqrt(d2[gi[i]])
            if reach[i]:
                assert d <= 1.0 + 1e-6 and np.array_equal(eb[i], targets[i])
            else:
                assert d > 1.0 - 1e-6 and abs(np.linalg.norm(eb[i].astype(np.float64) - ea[i]) - 1.0) < 1e-5
        assert gi[0] == 17 and reach[0] == 1
        if world == 1:
            first = (gi.copy(), eb.copy(), reach.copy())
        else:  # the decision does not depend on how the tree is dealt out
            assert np.array_equal(gi, first[0]) and np.array_equal(eb, first[1]) and np.array_equal(reach, first[2])


def test_sharded_tree_planner_follows_the_device_planner(pv, c64):
    """distributed.ShardedTreePlanner (world size 1 here; worlds 2 and 3 run over gloo in the CPU suite, N GPUs in
    tools/multi_gpu_tree.py): nearest-node candidates, steering, batched motion validation and the RRT-Connect
    transitions as separate device steps reproduce the one-kernel planner query by query."""
    from rbe550_final_project_b200.distributed import ShardedTreePlanner
    wall = sc.make_obb((0.55, 0.0, 0.35), (0.5, 0.04, 0.7))
    snap = sc.SceneSnapshot(obb=np.array([wall], dtype=np.float32), names=["wall"], entity_idx=[1])
    pv.set_scene(snap)
    pv.set_attached(-1)
    pv.set_flags(True, False)
    cand = random_configs(400, 55)
    valid = cand[unpack_bits(pv.check_states_host(cand), len(cand))]
    nq = 96
    starts, goals = valid[:nq].copy(), valid[nq:2 * nq].copy()
    starts[5] = pm.Q_LOWER - 0.1  # an out-of-bounds start: dropped at intake, like OMPL does
    kw = dict(max_iters=400, max_path=96, seed=21)
    ref = pv.rrtc_batch(starts, goals, max_nodes=1024, replicas=1, shortcut_passes=0, check_endpoints=True, **kw)
    assert (ref[2] > 1).sum() > 10, "some searches must really grow trees"
    pl = ShardedTreePlanner(pv, max_nodes=1024)
    paths, iters, status = pl.solve(starts, goals, check_endpoints=True, **kw)
    assert status[5] >= pl.BADEND and len(paths[5]) == 0 and ref[1][5] == 0
    same = 0
    for k in range(nq):
        if len(paths[k]) == ref[1][k] and np.array_equal(paths[k], ref[0][k, : ref[1][k]]) and iters[k] == ref[2][k]:
            same += 1
        if len(paths[k]):
            assert np.array_equal(paths[k][0], starts[k]) and np.array_equal(paths[k][-1], goals[k])
            _path_ok(pv, c64, snap, paths[k])
    # the batched edge kernel evaluates sin/cos in hardware, the in-kernel validator does not: a verdict inside the contact
    # band may differ, everything else is decision for decision the same
    assert same >= nq - 2, f"{same}/{nq} queries identical to pv_rrtc_batch"
    # split invariance through query_offset, like the one-kernel planner
    part = ShardedTreePlanner(pv, max_nodes=1024).solve(starts[40:], goals[40:], check_endpoints=True, query_offset=40, **kw)
    assert all(np.array_equal(a, b) for a, b in zip(part[0], paths[40:])) and np.array_equal(part[1], iters[40:])
