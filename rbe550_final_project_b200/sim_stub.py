"""Headless stand-ins for the Genesis objects the planner AND the reference's motion primitives touch.

Genesis is not installable here, so tests and bench.py drive `PlannerInterface` -- and the reference's own
`MotionPrimitiveExecutor` (tests/test_gpu_dropin.py) -- with these duck-typed objects.  They expose exactly the
attributes the reference reads (motion_primitives.py uses `robot.get_qpos / get_link /
inverse_kinematics / control_dofs_position`, `scene.step`, `block.get_pos`, `link.get_pos`): `robot.n_qs`, `n_dofs`,
`_solver.n_envs`, `q_limit`, `get_qpos()`, `set_qpos()`, `get_pos()`; `scene.entities`; per block
`idx`, `morph.size`, `get_pos()`, `get_quat()`.

Physics is KINEMATIC: `scene.step()` moves the robot to its commanded joint positions, fingers closing on a block stop
at its faces, and a block held between the closed fingers rides rigidly with the hand until they open -- enough for
pick / place / stack choreography to run end to end; nothing falls, slips or gets pushed.
"""
from __future__ import annotations

from types import SimpleNamespace
from typing import Dict, Tuple

import numpy as np

from . import panda_model as pm
from . import scenes as sc
from .robot_adapter import RobotAdapter


class StubEntity:
    def __init__(self, idx: int, name: str, pos, quat=(1.0, 0.0, 0.0, 0.0), size=None):
        self.idx = idx
        self.name = name
        self._pos = np.asarray(pos, dtype=np.float64)
        self._quat = np.asarray(quat, dtype=np.float64)
        self.morph = SimpleNamespace(size=tuple(size) if size is not None else None)

    def get_pos(self):
        return self._pos.copy()

    def get_quat(self):
        return self._quat.copy()

    def set_pos(self, pos):
        self._pos = np.asarray(pos, dtype=np.float64)

    def set_quat(self, quat):
        self._quat = np.asarray(quat, dtype=np.float64)


class StubPanda(StubEntity):
    n_qs = pm.N_Q
    n_dofs = pm.N_Q

    def __init__(self, idx: int):
        super().__init__(idx, "panda", pm.BASE_LIFT)
        self._solver = SimpleNamespace(n_envs=0)
        self.q_limit = np.stack([pm.Q_LOWER, pm.Q_UPPER])
        self._q = pm.Q_SCENE_INIT.copy()
        self._target = None
        self.ik_mode = "kinematic"
        self._validity = None
        self._scene = None
        self.set_qpos_calls = 0

    def get_qpos(self):
        return self._q.copy()

    # --- kinematics through the CUDA library (only when a PandaValidity handle has been attached) ----------
    def attach_validity(self, validity):
        self._validity = validity

    def get_link(self, name: str):
        idx = pm.LINK_NAMES.index(name)
        robot = self

        class _Link:
            def __init__(self):
                self.name = name
                self.idx = idx

            def get_pos(self):
                import torch
                pose = robot._validity.fk(torch.as_tensor(robot._q[None], dtype=torch.float32, device=robot._validity.device))
                return pose[0, idx, 0:3].cpu().numpy().astype(np.float64)

        return _Link()

    def inverse_kinematics(self, link=None, pos=None, quat=None, **kw):
        """robot.inverse_kinematics(link=hand, pos=, quat=) (motion_primitives.py:131-134) on the GPU: qpos (9,) near the
        current one, or None.

        ik_mode "kinematic" (default) is what Genesis gives the reference: a solution inside the joint limits, chosen
        with NO regard for the scene -- the reference asks for poses that only make sense that way, e.g. the place pose
        of put_down while the "attached" block is still an obstacle at its old place.  (The candidates
        are still filtered for SELF-collision, which no caller could use.)  ik_mode "collision_aware" filters them by the
        full validity rule of the current world: live block poses, the held block attached (planning.py:216-230)."""
        if getattr(link, "name", "hand") != "hand":
            raise NotImplementedError("only the hand link is supported")
        v = self._validity
        held = getattr(self._scene, "held", None)
        prev_scene = v.scene
        if self.ik_mode == "kinematic":
            v.set_scene(sc.SceneSnapshot(obb=np.zeros((0, 16), np.float32), table_z=-10.0, base=tuple(self._pos)))
        elif self._scene is not None:
            snap = sc.snapshot_from_sim(self._scene, self)
            v.set_scene(snap)
            if held is not None:
                v.set_attached(snap.index_of_entity(held[0].idx))
        try:
            return v.ik(np.asarray(pos, dtype=np.float64), np.asarray(quat, dtype=np.float64), self._q, **kw)
        finally:
            if prev_scene is not None:
                v.set_scene(prev_scene)  # (also leaves nothing attached)
            else:
                v.set_attached(-1)

    def set_qpos(self, q):
        self.set_qpos_calls += 1
        self._q = np.asarray(q.detach().cpu().numpy() if hasattr(q, "detach") else q, dtype=np.float64).copy()

    # --- position control (motion_primitives.py:171, 216, 245, ...) ---------------------------------------
    def control_dofs_position(self, q, dofs_idx_local=None):
        q = np.asarray(q.detach().cpu().numpy() if hasattr(q, "detach") else q, dtype=np.float64).reshape(-1)
        target = self._q.copy() if self._target is None else self._target.copy()
        if dofs_idx_local is None:
            target[: len(q)] = q
        else:
            target[np.asarray(dofs_idx_local, dtype=int)] = q
        self._target = target

    def hand_pose(self):
        """(position (3,), rotation (3,3)) of the hand link from the FK kernel."""
        import torch
        if self._validity.scene is None and self._scene is not None:  # FK places the chain on the scene's base pose
            self._validity.set_scene(sc.snapshot_from_sim(self._scene, self))
        pose = self._validity.fk(torch.as_tensor(self._q[None], dtype=torch.float32, device=self._validity.device))
        pose = pose[0, pm.LINK_NAMES.index("hand")].double().cpu().numpy()
        return pose[:3], pose[3:].reshape(3, 3)


class StubScene:
    """Entities + a kinematic step: robot to its commanded position, held block with the hand."""

    # a block whose centre, in the hand frame, is within these distances of (0, 0, PAD_Z) sits between the pads; PAD_Z is
    # the grasp offset of motion_primitives.py:27 (the hand 0.12 m above the block centre: the finger tips, 0.112 m down
    # the hand's z axis, then hold the top centimetre of the block)
    GRASP_ZONE = (0.03, 0.03, 0.035)
    PAD_Z = 0.12

    def __init__(self):
        self.entities = []
        self.robot = None
        self.steps = 0
        self.held = None           # (block entity, R hand-from-block, t hand-from-block)

    def add(self, ent):
        self.entities.append(ent)
        if isinstance(ent, StubPanda):
            self.robot = ent
        return ent

    def _blocks(self):
        return [e for e in self.entities if e is not self.robot and e.morph.size is not None]

    def step(self):
        self.steps += 1
        rob = self.robot
        if rob is None or rob._target is None:
            return
        q = rob._target.copy()
        q[:7] = np.clip(q[:7], pm.Q_LOWER[:7], pm.Q_UPPER[:7])
        q[7:] = np.clip(q[7:], pm.Q_LOWER[7:], pm.Q_UPPER[7:])
        opening = q[7] + q[8]
        if getattr(rob, "_validity", None) is None:
            rob._q = q
            return
        if self.held is not None:
            blk, R_hb, t_hb = self.held
            width = float(blk.morph.size[1])
            if opening > width + 2e-3:
                self.held = None       # released: the block stays where it is
            else:
                q[7:] = np.maximum(q[7:], 0.5 * width - 5e-4)  # the pads stop at the block's faces
        rob._q = q
        p, R = rob.hand_pose()
        if self.held is not None:
            blk, R_hb, t_hb = self.held
            blk.set_pos(p + R @ t_hb)
            blk.set_quat(_mat_to_quat(R @ R_hb))
            return
        # fingers closing on a block that sits between the pads: it is grasped from now on
        for blk in self._blocks():
            width = float(blk.morph.size[1])
            if opening > width + 2e-3:
                continue
            loc = R.T @ (blk.get_pos() - p)
            if abs(loc[0]) < self.GRASP_ZONE[0] and abs(loc[1]) < self.GRASP_ZONE[1] and abs(loc[2] - self.PAD_Z) < self.GRASP_ZONE[2]:
                Rb = sc.quat_wxyz_to_mat(blk.get_quat())
                self.held = (blk, R.T @ Rb, loc)
                rob._q[7:] = np.maximum(rob._q[7:], 0.5 * width - 5e-4)
                break


def _mat_to_quat(R) -> np.ndarray:
    """Rotation matrix -> quaternion wxyz (Shepperd)."""
    t = np.trace(R)
    if t > 0:
        s = np.sqrt(t + 1.0) * 2
        q = [0.25 * s, (R[2, 1] - R[1, 2]) / s, (R[0, 2] - R[2, 0]) / s, (R[1, 0] - R[0, 1]) / s]
    elif R[0, 0] > R[1, 1] and R[0, 0] > R[2, 2]:
        s = np.sqrt(1.0 + R[0, 0] - R[1, 1] - R[2, 2]) * 2
        q = [(R[2, 1] - R[1, 2]) / s, 0.25 * s, (R[0, 1] + R[1, 0]) / s, (R[0, 2] + R[2, 0]) / s]
    elif R[1, 1] > R[2, 2]:
        s = np.sqrt(1.0 + R[1, 1] - R[0, 0] - R[2, 2]) * 2
        q = [(R[0, 2] - R[2, 0]) / s, (R[0, 1] + R[1, 0]) / s, 0.25 * s, (R[1, 2] + R[2, 1]) / s]
    else:
        s = np.sqrt(1.0 + R[2, 2] - R[0, 0] - R[1, 1]) * 2
        q = [(R[1, 0] - R[0, 1]) / s, (R[0, 2] + R[2, 0]) / s, (R[1, 2] + R[2, 1]) / s, 0.25 * s]
    return np.asarray(q, dtype=np.float64)


def yaw_quat(deg: float):
    a = np.radians(deg) / 2.0
    return (float(np.cos(a)), 0.0, 0.0, float(np.sin(a)))


def scene_from_snapshot(snap: sc.SceneSnapshot) -> Tuple[StubScene, RobotAdapter, Dict[str, StubEntity]]:
    """Entity order of the reference factories: plane, blocks, robot last (scenes.py:49-85)."""
    scene = StubScene()
    scene.add(StubEntity(0, "plane", (0, 0, 0)))
    blocks: Dict[str, StubEntity] = {}
    for k in range(snap.n_obb):
        o = snap.obb[k].astype(np.float64)
        R = o[6:15].reshape(3, 3)
        yaw = np.degrees(np.arctan2(R[1, 0], R[0, 0]))
        ent = StubEntity(k + 1, snap.names[k], o[0:3], yaw_quat(yaw), size=2.0 * o[3:6])
        blocks[snap.names[k]] = scene.add(ent)
    robot = scene.add(StubPanda(snap.n_obb + 1))
    robot._scene = scene
    return scene, RobotAdapter(robot, scene), blocks


def create_scene(name: str = "goal1_scattered", **kw):
    """Headless mirror of the reference scene factories (scenes.py:41-373): (scene, franka, blocks_state)."""
    return scene_from_snapshot(sc.FIXTURES[name](**kw))
