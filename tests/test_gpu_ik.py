"""GPU tests of the batched, collision-aware IK (next-row component 8f-2) and of the motion-primitive call
sequence (motion_primitives.py:256-420) driven through the drop-in planner.  IK has no unique answer, so parity is
by property: the oracle's FK of the returned configuration reaches the requested pose, the configuration is inside
the joint limits and collision-free in the oracle."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import panda_oracle as po
from rbe550_final_project_b200 import panda_model as pm
from rbe550_final_project_b200 import scenes as sc
from rbe550_final_project_b200.planning import PlannerInterface
from rbe550_final_project_b200.sim_stub import create_scene
from rbe550_final_project_b200.validity import unpack_bits

pytestmark = pytest.mark.gpu
GOALS = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "goal_configs.json")))
GRASP_QUAT = np.array([0.0, 1.0, 0.0, 0.0])  # motion_primitives.py:39


def _pose_error(q, pos, quat):
    R, p = po.fk(np.asarray(q, dtype=np.float64)[None])
    Rt = sc.quat_wxyz_to_mat(quat)
    rel = R[0, 8].T @ Rt
    ang = np.arccos(np.clip((np.trace(rel) - 1) / 2, -1, 1))
    return float(np.linalg.norm(p[0, 8] - pos)), float(ang)


def test_ik_reaches_reference_grasp_poses(pv, model):
    for scene_name, cases in GOALS.items():
        if scene_name == "safe_home":
            continue
        snap = sc.FIXTURES[scene_name]()
        pv.set_scene(snap)
        pv.set_attached(-1)
        names = list(cases)
        pos = np.array([cases[n]["hand_pos"] for n in names])
        quat = np.tile(GRASP_QUAT, (len(names), 1))
        q, ok, err = pv.ik_batch(pos, quat, pm.Q_SAFE_HOME, n_seeds=128)
        assert ok.all(), (scene_name, names, ok)
        for k in range(len(names)):
            dp, da = _pose_error(q[k], pos[k], GRASP_QUAT)
            assert dp < 2e-4 and da < 2e-3, (names[k], dp, da)
            assert po.in_bounds(q[k].astype(np.float64)[None], model)[0]
            assert po.state_margin(q[k].astype(np.float64)[None], snap.as_oracle_scene(), model)[0] > -1e-4
            assert np.allclose(q[k, 7:], pm.Q_SAFE_HOME[7:])
        # the returned solution is the valid one closest to the initial configuration among the seeds: it must be
        # at least as close as the golden configuration found offline from the same start
        for k, n in enumerate(names):
            d_gpu = np.linalg.norm(q[k, :7] - pm.Q_SAFE_HOME[:7])
            d_gold = np.linalg.norm(np.array(cases[n]["q"])[:7] - pm.Q_SAFE_HOME[:7])
            assert d_gpu < d_gold + 0.5


def test_ik_rejects_unreachable_and_colliding_poses(pv):
    pv.set_scene(sc.goal1_scattered())
    pos = np.array([[1.5, 0.0, 0.5],      # out of reach
                    [0.65, 0.0, 0.05],    # hand 5 cm above the table, pointing down: fingers in the table / block r
                    [0.5, 0.0, 0.3]])     # fine
    q, ok, err = pv.ik_batch(pos, np.tile(GRASP_QUAT, (3, 1)), pm.Q_SAFE_HOME, n_seeds=128)
    assert list(ok) == [False, False, True]


def test_ik_is_deterministic_and_batched(pv):
    pv.set_scene(sc.goal3_tower())
    rng = np.random.default_rng(3)
    n = 512
    pos = np.stack([rng.uniform(0.3, 0.7, n), rng.uniform(-0.4, 0.4, n), rng.uniform(0.15, 0.6, n)], axis=1)
    quat = np.tile(GRASP_QUAT, (n, 1))
    a = pv.ik_batch(pos, quat, pm.Q_SAFE_HOME, n_seeds=64, seed=5)
    b = pv.ik_batch(pos, quat, pm.Q_SAFE_HOME, n_seeds=64, seed=5)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    assert a[1].mean() > 0.9
    valid = unpack_bits(pv.check_states_host(a[0][a[1]]), int(a[1].sum()))
    assert valid.all()
    for k in np.nonzero(a[1])[0][:50]:
        dp, da = _pose_error(a[0][k], pos[k], GRASP_QUAT)
        assert dp < 2e-4 and da < 2e-3


def test_pick_and_place_call_sequence(pv, c64):
    """The call pattern of MotionPrimitiveExecutor.pick_up / put_down (motion_primitives.py:256-302, 356-420):
    IK for the approach pose -> plan_path -> IK for the grasp pose -> plan_path, then with the block attached
    IK + plan_path to the place-approach pose.  Every returned path must be valid in the oracle."""
    scene, franka, blocks = create_scene("goal1_scattered")
    franka.raw.attach_validity(pv)
    franka.set_qpos(pm.Q_SAFE_HOME)
    planner = PlannerInterface(franka, scene, validity=pv)
    planner.refresh_scene()
    hand = franka.get_link("hand")
    snap = sc.goal1_scattered()

    def plan_to(q_goal, attached=None, att_idx=-1):
        path = planner.plan_path(qpos_goal=q_goal, num_waypoints=150, attached_object=attached, timeout=10.0)
        assert len(path) == 150
        arr = np.stack([w.numpy() for w in path]).astype(np.float64)
        m = c64.edge_margin(arr[:-1], arr[1:], snap.as_oracle_scene(), n_steps=0, attached=att_idx)
        assert (m > -1e-4).all()
        franka.set_qpos(arr[-1])  # "execute"

    for key in ("r", "c"):
        franka.set_qpos(pm.Q_SAFE_HOME)
        center = np.array(blocks[key].get_pos())
        # pick_up: approach = block top + MIN_APPROACH_HEIGHT (0.18), grasp = centre + grasp_offset (0.12)
        q_app = franka.inverse_kinematics(link=hand, pos=center + [0, 0, 0.02 + 0.18], quat=GRASP_QUAT)
        assert q_app is not None
        plan_to(q_app)
        assert np.allclose(hand.get_pos(), center + [0, 0, 0.20], atol=3e-4)
        q_grasp = franka.inverse_kinematics(link=hand, pos=center + [0, 0, 0.12], quat=GRASP_QUAT)
        assert q_grasp is not None
        plan_to(q_grasp)
        # close the gripper onto the block (0.04 cube -> fingers at 0.02 touch it): only valid when attached
        q_hold = franka.get_qpos()
        q_hold[7:] = 0.0195
        franka.set_qpos(q_hold)
        att = snap.index_of_entity(blocks[key].idx)
        # put_down: approach pose 0.15 above the place pose (0.50, 0.0), block attached (stays a static obstacle)
        place = np.array([0.50, 0.0, 0.02 + 0.12 + 0.15])
        pv.set_attached(att)
        q_place = franka.inverse_kinematics(link=hand, pos=place, quat=GRASP_QUAT)
        assert q_place is not None
        q_place[7:] = 0.0195
        plan_to(q_place, attached=blocks[key], att_idx=att)
        pv.set_attached(-1)
