"""RobotAdapter -- the duck-typed robot the planner receives (mirrors robot_adapter.py:20-72).

A transparent proxy: unknown attributes are forwarded to the wrapped entity (robot_adapter.py:31-37), a
few calls are forwarded explicitly so their names are guaranteed to exist (robot_adapter.py:41-67), and
`.raw` exposes the wrapped object (robot_adapter.py:70-72).  Written from the interface, not copied.
"""
from typing import Any


class RobotAdapter:
    _FORWARDED = (
        "get_pos", "set_pos", "get_qpos", "set_qpos", "control_dofs_position", "control_dofs_force",
        "get_link", "inverse_kinematics", "detect_collision",
    )

    def __init__(self, robot: Any, scene: Any = None):
        self.robot = robot
        self.scene = scene

    def __getattr__(self, name: str) -> Any:
        # only reached for names not found on the adapter itself
        return getattr(self.__dict__["robot"], name)

    @property
    def raw(self) -> Any:
        return self.robot


def _make_forward(name):
    def fwd(self, *args, **kwargs):
        return getattr(self.robot, name)(*args, **kwargs)
    fwd.__name__ = name
    return fwd


for _n in RobotAdapter._FORWARDED:
    setattr(RobotAdapter, _n, _make_forward(_n))
