"""Drop-in for the reference's `planning.py`: `PlannerInterface(robot, scene).plan_path(...)`.

Same constructor, same `plan_path` signature, same return convention and the same soft/hard error
behaviour as planning.py:24-242 -- but every validity query (planning.py:209-219) is answered by the
sm_100a kernels behind `libpanda_validity.so`, and the whole of plan_path -- intake checks, RRT-Connect solve,
simplifySolution, interpolate, and a dense validation of the waypoints handed back -- is ONE call into the C-ABI
(`pv_plan_path`) instead of one Python callback per sampled state.  `motion_primitives.py` uses
it unchanged (`from planning import PlannerInterface`, motion_primitives.py:9; see INTEGRATION.md and
tests/test_gpu_dropin.py, which drives the reference's own MotionPrimitiveExecutor through it).

Differences that are deliberate and visible:
  * the simulated robot is never moved during planning (the reference poses it for every state and
    restores it at planning.py:205); `robot.set_qpos(qpos_cur)` is still issued at exit when the robot
    offers it, so callers relying on that side effect see the same final state;
  * RRTConnect (the default every caller uses) and RRT run on the device; the other six names of
    planning.py:108-117 (PRM, RRTstar, EST, FMT, BITstar, ABITstar) pass the name guard and are answered by the
    device RRTConnect + simplification with a logged warning -- `strict_planners=True` raises instead;
  * the returned waypoints are validated more densely than OMPL validates them (every waypoint and every motion
    between consecutive waypoints); a path that grazes a contact between the planner's own samples is re-planned
    instead of being returned (`validate_waypoints=False` restores the reference's behaviour);
  * the waypoints are the rows of one (n, 9) fp32 CPU tensor, handed out as a sequence (`Waypoints`) that creates the
    per-row tensors when they are read -- `len`, truth value, iteration, indexing and `np.array(path[-1])` behave as
    for the reference's list of tensors (motion_primitives.py:146-176); `[]` is returned on failure;
  * there is no CPU fallback: without the CUDA library / a B200 the constructor raises.
"""
from __future__ import annotations

import collections.abc
import logging
import time
from typing import Any, List, Optional, Sequence

import numpy as np
import torch

from . import panda_model as pm
from .robot_adapter import RobotAdapter
from .scenes import SceneSnapshot, snapshot_from_sim
from .validity import PandaValidity, decode_culprit

logger = logging.getLogger("panda_validity.planning")

SUPPORTED_PLANNERS = ["PRM", "RRT", "RRTConnect", "RRTstar", "EST", "FMT", "BITstar", "ABITstar"]  # planning.py:108-117
DEVICE_PLANNERS = ["RRTConnect", "RRT"]


class Waypoints(collections.abc.Sequence):
    """The waypoints of one plan: rows of ONE (n, 9) fp32 CPU tensor (`.tensor`; `.array` is the numpy view).
    Behaves like the reference's list of n length-9 tensors (planning.py:200, 232-242) for everything its callers do
    (motion_primitives.py:146-176: truth value, `for wp in path`, `wp.cpu().numpy()`, `path[-1]`), but the n small
    tensor objects are only created when they are read: building 150 of them costs more than planning does."""

    __slots__ = ("array", "tensor", "_rows")

    def __init__(self, array: np.ndarray):
        self.array = np.ascontiguousarray(array, dtype=np.float32).reshape(-1, pm.N_Q)
        self.tensor = torch.from_numpy(self.array)
        self._rows = None

    def _materialise(self):
        if self._rows is None:
            self._rows = list(self.tensor.unbind(0))
        return self._rows

    def __len__(self) -> int:
        return self.array.shape[0]

    def __getitem__(self, i):
        if isinstance(i, slice):
            return self._materialise()[i]
        if self._rows is not None:
            return self._rows[i]
        return self.tensor[i]

    def __iter__(self):
        return iter(self._materialise())

    def __eq__(self, other):
        if isinstance(other, (list, tuple, Waypoints)):
            return len(self) == len(other) and all(torch.equal(a, torch.as_tensor(b)) for a, b in zip(self, other))
        return NotImplemented

    __hash__ = None

    def __repr__(self) -> str:
        return f"Waypoints(n={len(self)})"


_EPS = float(np.finfo(np.float64).eps)
_LOWER = tuple(float(v) for v in pm.Q_LOWER)
_UPPER = tuple(float(v) for v in pm.Q_UPPER)


def _satisfies_bounds(q: np.ndarray) -> bool:
    """ob.RealVectorBounds check of planning.py:165-173 (satisfiesBounds: eps slack on both sides); plain floats, this
    sits inside the plan time."""
    for v, lo, hi in zip(q.tolist(), _LOWER, _UPPER):
        if not (v - _EPS <= hi and v + _EPS >= lo):
            return False
    return True


class PlanningError(Exception):
    """Raised where the reference calls gs.raise_exception (planning.py:106,119,122,125,135)."""


def tensor_to_array(x) -> np.ndarray:
    """genesis.utils.misc.tensor_to_array as used at planning.py:131-132."""
    if torch.is_tensor(x):
        return x.detach().cpu().numpy()
    return np.asarray(x)


def _ensure_adapter(robot: Any, scene: Any) -> RobotAdapter:
    # planning.py:14-22
    return robot if isinstance(robot, RobotAdapter) else RobotAdapter(robot, scene)


class StateValidityChecker:
    """Callable with the shape OMPL's `ob.StateValidityCheckerFn` expects (planning.py:155): `state[i]`
    indexable for i < 9 -> bool.  One state per call goes through the host entry point of the C-ABI;
    batches should use `PandaValidity.check_states` directly."""

    def __init__(self, validity: PandaValidity, n_qs: int = pm.N_Q):
        self.validity = validity
        self.n_qs = n_qs
        self.calls = 0

    def __call__(self, state) -> bool:
        self.calls += 1
        q = np.array([float(state[i]) for i in range(self.n_qs)], dtype=np.float32)
        return self.validity.is_state_valid(q)


class PlannerInterface:
    def __init__(self, robot: Any, scene: Any, device: int = 0, validity: Optional[PandaValidity] = None,
                 carry_attached: bool = False, strict_planners: bool = False, validate_waypoints: bool = True):
        self.robot = _ensure_adapter(robot, scene)
        self.scene = scene
        self.attached_object = None
        # False = the reference's rule (planning.py:216-230): the grasped block stays an obstacle where it is and only
        # hand / finger contacts with it are forgiven.  True = the block rides on the hand:
        # its grasp pose is taken from the robot configuration and block pose at the moment plan_path is called.
        self.carry_attached = bool(carry_attached)
        # The six planner names of planning.py:108-117 without a device kernel (PRM, RRTstar, EST, FMT, BITstar, ABITstar)
        # are answered by the device RRT-Connect + simplification with a warning, so that a caller passing them keeps
        # working; strict_planners=True raises instead.  (Every caller in the reference uses the default.)
        self.strict_planners = bool(strict_planners)
        # dense validation of the returned waypoints with fallback / re-planning (pv_plan_path); False = hand back
        # whatever the planner found, as OMPL does
        self.validate_waypoints = bool(validate_waypoints)
        self.validity = validity if validity is not None else PandaValidity(device)
        self._snapshot: Optional[SceneSnapshot] = None
        self._q_grasp = None
        self.rng_seed = 1
        self.replicas = 32          # independent device searches per plan; smallest (iterations, replica id) wins
        self.last_stats: dict = {}

    # ---- scene --------------------------------------------------------------------------------------
    def refresh_scene(self) -> SceneSnapshot:
        """Freeze the block poses for this plan (they do not move inside plan_path).  The handle only
        re-derives its scene tables when the snapshot differs from the one it already holds."""
        snap = self.scene if isinstance(self.scene, SceneSnapshot) else snapshot_from_sim(self.scene, self.robot)
        self.validity.set_scene(snap)
        self._snapshot = snap
        return snap

    def _attached_index(self, attached_object) -> int:
        if attached_object is None:
            return -1
        if isinstance(attached_object, (int, np.integer)):
            return int(attached_object)
        idx = getattr(attached_object, "idx", None)
        if idx is None:
            raise PlanningError("attached_object has no .idx (planning.py:226)")
        return self._snapshot.index_of_entity(int(idx))

    def _apply_attached(self, attached_object, q_grasp=None):
        k = self._attached_index(attached_object)
        if self.carry_attached and k >= 0 and q_grasp is not None:
            self.validity.set_carried(k, q_grasp=np.asarray(tensor_to_array(q_grasp), dtype=np.float32))
        else:
            self.validity.set_attached(k)

    def _with_attached(self, fn):
        """Run fn() with THIS planner's attached object (planning.py:153 keeps it after plan_path, and the callback
        of planning.py:216-219 reads it) applied to the handle, then leave the handle with nothing attached: the
        handle may be shared, and whoever uses it next must not inherit this planner's grasp."""
        if self._snapshot is None:
            self.refresh_scene()
        if self.attached_object is None:
            return fn()
        self._apply_attached(self.attached_object, self._q_grasp)
        try:
            return fn()
        finally:
            self.validity.set_attached(-1)

    # ---- diagnostics (planning.py:32-57) ----------------------------------------------------------------
    def diagnose_bounds_violation(self, state, lower=None, upper=None):
        lower = pm.Q_LOWER if lower is None else lower
        upper = pm.Q_UPPER if upper is None else upper
        violated = [(i, float(state[i]), float(lower[i]), float(upper[i]))
                    for i in range(self.robot.n_qs) if state[i] < lower[i] or state[i] > upper[i]]
        logger.warning(f"State violates bounds on joints: {violated}")
        return violated

    def diagnose_valid_violation(self, state):
        q = np.array([float(state[i]) for i in range(pm.N_Q)], dtype=np.float32)

        def run():
            m, cu = self.validity.state_margins(torch.as_tensor(q[None], device=self.validity.device), want_culprit=True)
            return m, cu, self.validity.contacts(torch.as_tensor(q[None], device=self.validity.device))[0]

        m, cu, pairs = self._with_attached(run)
        culprit = decode_culprit(int(cu[0].item()))
        bad_links = sorted({name for pair in pairs for name in pair})
        # planning.py:49-57 prints the link names of every contact pair
        logger.warning(f"State causes collisions between links: {bad_links}; deepest: {culprit} "
                       f"(clearance {float(m[0].item()):.4f} m)")
        return culprit

    # ---- the validity callback (planning.py:209-219) ---------------------------------------------------------
    def _is_ompl_state_valid(self, state) -> bool:
        q = np.array([float(state[i]) for i in range(pm.N_Q)], dtype=np.float32)
        return self._with_attached(lambda: self.validity.is_state_valid(q))

    def state_validity_checker(self) -> StateValidityChecker:
        if self._snapshot is None:
            self.refresh_scene()
        return StateValidityChecker(self.validity, self.robot.n_qs)

    def check_motion(self, qa: Sequence[float], qb: Sequence[float]) -> bool:
        """si.checkMotion(a, b) with the DiscreteMotionValidator defaults."""
        w = self._with_attached(lambda: self.validity.check_edges_host(
            np.asarray(qa, np.float32)[None], np.asarray(qb, np.float32)[None], n_steps=0))
        return bool(w[0] & 1)

    def simplify_path(self, path: np.ndarray, seed: int = 1) -> np.ndarray:
        """`ss.simplifySolution()` (planning.py:196) on a given vertex list: OMPL's simplifyMax passes (partial
        shortcuts, B-spline smoothing, vertex reduction), each validated as one batch by the edge kernel
        (pv_simplify_path).  plan_path does this inside pv_plan_path; this entry exists for callers that bring their
        own path."""
        return self._with_attached(lambda: self.validity.simplify_path(np.asarray(path, dtype=np.float64), seed=seed))

    def validate_trajectory(self, waypoints, attached_object=None, q_grasp=None) -> np.ndarray:
        """Swept validation of an executed joint trajectory (next-row component 8f-4): the reference plays back the
        150 planned waypoints and many un-planned joint-space lerps (motion_primitives.py:163-173, 294-299, 404-409)
        with no collision check.  Returns one bool per segment (waypoint k -> k+1), each segment discretised at the
        motion-validity resolution; a motion check assumes its start state valid (as OMPL's does), so the FIRST
        waypoint is checked as a state of its own and an invalid one fails segment 0.  Dense waypoints make this
        check FINER than the planner's validator (one state every 1 % of the space extent)."""
        if self._snapshot is None:
            self.refresh_scene()
        if isinstance(waypoints, Waypoints):
            pts = waypoints.array
        else:
            pts = np.stack([tensor_to_array(w) for w in waypoints]).astype(np.float32)
        if len(pts) < 2:
            return np.ones(0, dtype=bool)
        # carry mode: the block is where the snapshot saw it when the robot is at q_grasp (default: the first waypoint)
        self._apply_attached(attached_object, pts[0] if q_grasp is None else q_grasp)
        try:
            a = np.concatenate([pts[:1], pts[:-1]])  # motion 0 is (w0 -> w0): the first state itself
            bits = self.validity.check_edges_host(a, pts, n_steps=0)
        finally:
            self.validity.set_attached(-1)
        ok = ((bits[:, None] >> np.arange(32, dtype=np.uint32)) & 1).ravel()[: len(pts)].astype(bool)
        seg = ok[1:].copy()
        seg[0] &= ok[0]
        return seg

    # ---- plan_path (planning.py:59-207) ---------------------------------------------------------------------
    def plan_path(self, qpos_goal, qpos_start=None, timeout=5.0, smooth_path=True, num_waypoints=100,
                  attached_object=None, planner="RRTConnect"):
        if planner not in SUPPORTED_PLANNERS:
            raise PlanningError(f"Planner {planner} is not supported. Supported planners: {SUPPORTED_PLANNERS}.")
        if planner not in DEVICE_PLANNERS:
            if self.strict_planners:
                raise PlanningError(f"Planner {planner} has no device implementation; available: {DEVICE_PLANNERS}.")
            logger.warning(f"Planner {planner} has no device implementation: using the device RRTConnect + simplification.")
            planner = "RRTConnect"
        solver = getattr(self.robot, "_solver", None)
        if solver is not None and getattr(solver, "n_envs", 0) > 0:
            raise PlanningError("Motion planning is not supported for batched envs (yet).")
        if self.robot.n_qs != self.robot.n_dofs:
            raise PlanningError("Motion planning is not yet supported for rigid entities with free joints.")

        t0 = time.perf_counter()
        qpos_cur = self.robot.get_qpos()
        if qpos_start is None:
            qpos_start = qpos_cur
        qpos_start = np.asarray(tensor_to_array(qpos_start), dtype=np.float64)
        qpos_goal = np.asarray(tensor_to_array(qpos_goal), dtype=np.float64)
        if qpos_start.shape != (self.robot.n_qs,) or qpos_goal.shape != (self.robot.n_qs,):
            raise PlanningError("Invalid shape for `qpos_start` or `qpos_goal`.")
        if self.robot.n_qs != pm.N_Q:
            raise PlanningError(f"the device kernels are specialised to the {pm.N_Q}-DoF Panda; robot has {self.robot.n_qs}")

        # joint limits come from the robot (planning.py:139-140); the kernels carry the frozen Panda limits
        lower, upper = pm.Q_LOWER, pm.Q_UPPER
        q_limit = getattr(self.robot, "q_limit", None)
        if q_limit is not None and not getattr(self, "_limits_checked", False):
            lo_r = np.asarray(tensor_to_array(q_limit[0]), dtype=float)
            hi_r = np.asarray(tensor_to_array(q_limit[1]), dtype=float)
            if np.abs(lo_r - lower).max() > 1e-4 or np.abs(hi_r - upper).max() > 1e-4:
                logger.warning("robot.q_limit differs from the frozen Panda limits the kernels sample in")
            self._limits_checked = True

        snap = self.refresh_scene()
        t_scene = time.perf_counter()
        self.attached_object = attached_object  # planning.py:153
        self._q_grasp = np.asarray(tensor_to_array(qpos_cur), dtype=np.float32) if attached_object is not None else None
        if attached_object is not None:
            self._apply_attached(attached_object, self._q_grasp)  # (refresh_scene left the handle with nothing attached)

        # diagnostics on start / goal (planning.py:163-183): log, keep going
        start_in = _satisfies_bounds(qpos_start)
        goal_in = _satisfies_bounds(qpos_goal)
        if not start_in:
            logger.warning("OMPL start state out of bounds")
            self.diagnose_bounds_violation(qpos_start, lower, upper)
        if not goal_in:
            logger.warning("OMPL goal state out of bounds")
            self.diagnose_bounds_violation(qpos_goal, lower, upper)
        waypoints: Any = []
        stats = {"solved": False, "iters": 0, "state_checks": 0, "attempts": 0, "n_obb": snap.n_obb}
        # Validity of start and goal (planning.py:175-183) is judged inside the solve kernel; OMPL drops
        # out-of-bounds / invalid starts and goals at intake -> "no solution" 
        t_call = t_done = time.perf_counter()
        try:
            if start_in and goal_in:
                wp, st = self.validity.plan_path(
                    qpos_start, qpos_goal, num_waypoints=int(num_waypoints) if num_waypoints is not None else 0,
                    smooth=bool(smooth_path), planner=planner, seed=self.rng_seed, replicas=self.replicas,
                    validate=self.validate_waypoints, timeout=float(timeout))
                t_done = time.perf_counter()
                self.rng_seed += 1
                stats.update(st)
                stats["state_checks"] = st["checks"]
                stats["solved"] = bool(st["solved"])
                if st["endpoint_status"] & 1:
                    logger.warning("OMPL start state invalid")
                    self.diagnose_valid_violation(qpos_start)
                if st["endpoint_status"] & 2:
                    logger.warning("OMPL goal state invalid")
                    self.diagnose_valid_violation(qpos_goal)
                if len(wp):
                    logger.info("Path solution found successfully.")
                    print("Number of waypoints in path:", len(wp))  # planning.py:199
                    waypoints = Waypoints(wp)
        finally:
            # the handle may be shared: whoever uses it next must not inherit this plan's grasp
            if attached_object is not None:
                stats["carried"] = self.validity.carried  # (R, t, allowance) of the carry mode, None otherwise
                self.validity.set_attached(-1)
        if not len(waypoints):
            logger.warning("Path planning failed. Returning empty path.")
        t_end = time.perf_counter()
        stats["plan_ms"] = (t_end - t0) * 1e3
        stats["ms_scene_snapshot"] = (t_scene - t0) * 1e3
        stats["ms_c_call"] = (t_done - t_call) * 1e3
        stats["ms_python_rest"] = stats["plan_ms"] - stats["ms_scene_snapshot"] - stats["ms_c_call"]
        self.last_stats = stats

        # restore original state (planning.py:205) -- a no-op for us, kept for side-effect parity
        if hasattr(self.robot, "set_qpos"):
            try:
                self.robot.set_qpos(qpos_cur)
            except Exception:  # a stub robot without a simulator behind it
                pass
        return waypoints

    # ---- batched front end (BASELINE config 4) -------------------------------------------------------------
    def plan_paths_batch(self, starts: np.ndarray, goals: np.ndarray, max_iters: int = 2000, max_nodes: int = 2048,
                         smooth_path: bool = True, seed: int = 1, replicas: int = 1, max_path: int = 128,
                         packed: bool = False):
        """Many independent (start, goal) queries in one call.  Returns (paths, lengths, iters, checks), or with
        packed=True (states, offsets, lengths, iters, checks)."""
        if self._snapshot is None:
            self.refresh_scene()
        return self.validity.rrtc_batch(starts, goals, max_iters=max_iters, max_nodes=max_nodes, max_path=max_path,
                                        seed=seed, replicas=replicas, shortcut_passes=2 if smooth_path else 0,
                                        packed=packed)

    # planning.py:232-242
    def _ompl_state_to_tensor(self, state) -> torch.Tensor:
        return torch.tensor([float(state[i]) for i in range(self.robot.n_qs)], dtype=torch.float32)

    def _ompl_states_to_tensor_list(self, states) -> List[torch.Tensor]:
        return [self._ompl_state_to_tensor(s) for s in states]
