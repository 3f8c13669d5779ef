"""Headless stand-ins for the Genesis objects the planner touches.

Genesis is not installable here, so tests and bench.py drive `PlannerInterface` with these duck-typed
objects.  They expose exactly the attributes the reference reads: `robot.n_qs`, `n_dofs`,
`_solver.n_envs`, `q_limit`, `get_qpos()`, `set_qpos()`, `get_pos()`; `scene.entities`; per block
`idx`, `morph.size`, `get_pos()`, `get_quat()`.  No physics: poses only change when set.
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
        """robot.inverse_kinematics(link=hand, pos=, quat=) (motion_primitives.py:131-134) on the GPU: returns a
        collision-free qpos (9,) near the current one, or None."""
        if getattr(link, "name", "hand") != "hand":
            raise NotImplementedError("only the hand link is supported")
        return self._validity.ik(np.asarray(pos, dtype=np.float64), np.asarray(quat, dtype=np.float64), self._q, **kw)

    def set_qpos(self, q):
        self.set_qpos_calls += 1
        self._q = np.asarray(q.detach().cpu().numpy() if hasattr(q, "detach") else q, dtype=np.float64).copy()


class StubScene:
    def __init__(self):
        self.entities = []

    def add(self, ent):
        self.entities.append(ent)
        return ent


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
    return scene, RobotAdapter(robot, scene), blocks


def create_scene(name: str = "goal1_scattered", **kw):
    """Headless mirror of the reference scene factories (scenes.py:41-373): (scene, franka, blocks_state)."""
    return scene_from_snapshot(sc.FIXTURES[name](**kw))
