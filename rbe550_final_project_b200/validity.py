"""PandaValidity: Python face of the C-ABI handle (one per GPU).

PyTorch is used only for device buffers and streams; every computation is a hand-written sm_100a
kernel reached through ctypes.  There is no CPU path: constructing a PandaValidity without the built
library or without a B200 raises.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence, Tuple

import numpy as np
import torch

from . import _cabi
from . import panda_model as pm
from .scenes import SceneSnapshot

FLAG_SELF = 1
FLAG_LIMITS = 2

CULPRIT_KINDS = {0: "none", 1: "table", 2: "scene_box", 3: "self", 4: "joint_limit"}


class PandaValidityError(RuntimeError):
    pass


def soa_from_aos(q: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """(n, 9) float32 -> planes A (n, 4), B (n, 4), q9 (n,) as the kernels read them."""
    q = q.to(torch.float32)
    return q[:, 0:4].contiguous(), q[:, 4:8].contiguous(), q[:, 8].contiguous()


class PandaValidity:
    def __init__(self, device: int = 0):
        self.lib = _cabi.load()
        self._h = C.c_void_p()
        rc = self.lib.pv_create(int(device), C.byref(self._h))
        if rc != 0:
            raise PandaValidityError(f"pv_create failed ({rc}): {self.lib.pv_last_error(None).decode()}")
        self.device = torch.device("cuda", int(device))
        self.scene: Optional[SceneSnapshot] = None
        self._scene_key = None
        self._scene_obj_key = None
        self.attached = -1
        self.carried = None
        self.flags = FLAG_SELF
        self._plan_key = None
        self._plan_prm = None

    # -- plumbing -------------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self.lib.pv_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc: int, what: str):
        if rc != 0:
            raise PandaValidityError(f"{what} failed ({rc}): {self.lib.pv_last_error(self._h).decode()}")

    def _stream(self, stream) -> C.c_void_p:
        if stream is None:
            stream = torch.cuda.current_stream(self.device)  # the handle's device, whatever torch's current one is
        return C.c_void_p(stream.cuda_stream)

    @property
    def launch_count(self) -> int:
        return int(self.lib.pv_launch_count(self._h))

    # -- configuration --------------------------------------------------------------------------------
    def set_scene(self, scene: SceneSnapshot):
        if scene is self.scene and not scene.obb.flags.writeable and \
                self._scene_obj_key == (id(scene.obb), scene.table_z, scene.base):
            # the very snapshot object the handle already holds, read-only (scenes.snapshot_from_sim hands the previous
            # one back when no pose has changed): nothing to compare
            if self.attached != -1 or self.carried is not None:
                self._ck(self.lib.pv_set_attached(self._h, -1), "pv_set_attached")
                self.attached, self.carried = -1, None
            return
        obb = np.ascontiguousarray(scene.obb, dtype=np.float32).reshape(-1, 16)
        # plan_path snapshots the scene on every call (planning.py reads the live poses each time); the handle only
        # re-derives its tables (reach masks, scene bounds) when the bytes differ from what it already holds
        key = (obb.tobytes(), float(scene.table_z), tuple(float(v) for v in scene.base))
        self._scene_obj_key = (id(scene.obb), scene.table_z, scene.base)
        if key != self._scene_key:
            base = (C.c_float * 3)(*key[2])
            self._ck(self.lib.pv_set_scene(self._h, obb.ctypes.data_as(C.POINTER(C.c_float)), obb.shape[0],
                                           key[1], base), "pv_set_scene")
            self._scene_key = key
        self._ck(self.lib.pv_set_attached(self._h, -1), "pv_set_attached")  # a new snapshot starts with nothing held
        self.scene = scene
        self.attached = -1
        self.carried = None

    def set_attached(self, obb_index: int):
        self._ck(self.lib.pv_set_attached(self._h, int(obb_index)), "pv_set_attached")
        self.attached = int(obb_index)
        self.carried = None

    def set_carried(self, obb_index: int, hand_from_box=None, q_grasp=None, contact_allowance: float = 1e-3):
        """Physically-correct alternative to `set_attached` (not what planning.py:216-230
        does): scene box `obb_index` rides rigidly on the hand.  Give its pose in the hand frame as `hand_from_box` =
        (R (3,3), t (3,)), or the configuration `q_grasp` the robot has while the box sits at its snapshot pose (the
        moment of the grasp, motion_primitives.py:367-376).  The box is shrunk by `contact_allowance` in its own tests
        so that resting contacts (block on the table, block on a block) do not read as collisions.  -1 leaves the
        mode.  Returns (R, t)."""
        k = int(obb_index)
        if k < 0:
            self._ck(self.lib.pv_set_carried(self._h, -1, None, 0.0), "pv_set_carried")
            self.attached, self.carried = -1, None
            return None
        if hand_from_box is None:
            if q_grasp is None or self.scene is None or not (0 <= k < self.scene.n_obb):
                raise PandaValidityError("set_carried needs hand_from_box or q_grasp (and a valid scene-box index)")
            pose = self.fk(torch.as_tensor(np.asarray(q_grasp, dtype=np.float32).reshape(1, 9)))[0, 8].double().cpu().numpy()
            ph, Rh = pose[:3], pose[3:].reshape(3, 3)
            rec = np.asarray(self.scene.obb, dtype=np.float64).reshape(-1, 16)[k]
            R, t = Rh.T @ rec[6:15].reshape(3, 3), Rh.T @ (rec[0:3] - ph)
        else:
            R, t = (np.asarray(v, dtype=np.float64) for v in hand_from_box)
        buf = np.ascontiguousarray(np.concatenate([R.reshape(9), t.reshape(3)]), dtype=np.float32)
        self._ck(self.lib.pv_set_carried(self._h, k, buf.ctypes.data_as(C.POINTER(C.c_float)), float(contact_allowance)),
                 "pv_set_carried")
        self.attached, self.carried = k, (R.reshape(3, 3).copy(), t.reshape(3).copy(), float(contact_allowance))
        return self.carried

    def set_flags(self, self_collision: bool = True, joint_limits: bool = False):
        """`joint_limits` is kept for compatibility and changes nothing: a state outside the joint limits is always
        invalid (the pruned self-collision model is certified inside the limits only, see include/panda_validity.h)."""
        self.flags = (FLAG_SELF if self_collision else 0) | (FLAG_LIMITS if joint_limits else 0)
        self._ck(self.lib.pv_set_flags(self._h, self.flags), "pv_set_flags")

    def set_gather(self, peer_ptrs_dev: int, n_peers: int, multicast_ptr: int, word_offset: int, word_capacity: int = 0):
        """Fused verdict gather (see pv_set_gather); n_peers = 0 switches it off."""
        self._ck(self.lib.pv_set_gather(self._h, C.c_void_p(peer_ptrs_dev or None), int(n_peers),
                                        C.c_void_p(multicast_ptr or None), int(word_offset), int(word_capacity)), "pv_set_gather")

    def set_culling(self, mode):
        """State-kernel variant: 0 brute force, 1 per-lane bounding-ball culling, 2 (default) = 1 with each block's
        share of the batch visited in order of (wrist distance from the scene, elbow angle, wrist flex) so that warps
        agree on which tests to skip; bit-identical verdicts."""
        self._ck(self.lib.pv_set_culling(self._h, int(mode)), "pv_set_culling")

    def set_launch_overlap(self, on: bool):
        """Back-to-back check_states launches overlap (programmatic dependent launch; default on).  See pv_set_launch_overlap."""
        self._ck(self.lib.pv_set_launch_overlap(self._h, 1 if on else 0), "pv_set_launch_overlap")

    # -- device-buffer calls --------------------------------------------------------------------------
    def _planes(self, q) -> Tuple[torch.Tensor, torch.Tensor, Optional[torch.Tensor], int]:
        if isinstance(q, (tuple, list)) and len(q) in (2, 3) and torch.is_tensor(q[0]) and q[0].dim() == 2 \
                and q[0].shape[1] == 4:
            A, B = q[0], q[1]
            q9 = q[2] if len(q) == 3 else None
        else:
            qt = torch.as_tensor(q, dtype=torch.float32, device=self.device)
            if qt.dim() == 1:
                qt = qt[None]
            A, B, q9 = soa_from_aos(qt)
        for t in (A, B) + ((q9,) if q9 is not None else ()):
            if t.device != self.device or t.dtype != torch.float32 or not t.is_contiguous():
                raise PandaValidityError("config planes must be contiguous float32 tensors on the handle's device")
        return A, B, q9, int(A.shape[0])

    def check_states(self, q, out: Optional[torch.Tensor] = None, stream=None) -> torch.Tensor:
        """Verdict bit words (int32 tensor viewed as uint32 bits) for n configurations."""
        A, B, q9, n = self._planes(q)
        words = (n + 31) // 32
        if out is None:
            out = torch.empty(words, dtype=torch.int32, device=self.device)
        self._ck(self.lib.pv_check_states(self._h, A.data_ptr(), B.data_ptr(), q9.data_ptr() if q9 is not None else None,
                                          n, out.data_ptr(), self._stream(stream)), "pv_check_states")
        return out

    def state_margins(self, q, want_culprit: bool = False, stream=None):
        A, B, q9, n = self._planes(q)
        m = torch.empty(n, dtype=torch.float32, device=self.device)
        cu = torch.empty(n, dtype=torch.int32, device=self.device) if want_culprit else None
        self._ck(self.lib.pv_state_margins(self._h, A.data_ptr(), B.data_ptr(), q9.data_ptr() if q9 is not None else None,
                                           n, m.data_ptr(), cu.data_ptr() if cu is not None else None,
                                           self._stream(stream)), "pv_state_margins")
        return (m, cu) if want_culprit else m

    def contacts(self, q, stream=None):
        """detect_collision-like report: for each state the set of colliding (link, other) name pairs, where other
        is a robot link, 'ground' or 'box<k>' (scene box index)."""
        A, B, q9, n = self._planes(q)
        codes = torch.empty((n, 32), dtype=torch.int32, device=self.device)
        count = torch.empty(n, dtype=torch.int32, device=self.device)
        self._ck(self.lib.pv_state_contacts(self._h, A.data_ptr(), B.data_ptr(), q9.data_ptr() if q9 is not None else None,
                                            n, codes.data_ptr(), count.data_ptr(), self._stream(stream)), "pv_state_contacts")
        codes, count = codes.cpu().numpy(), count.cpu().numpy()
        out = []
        for i in range(n):
            pairs = set()
            for c in codes[i, : min(int(count[i]), 32)]:
                pairs.add(decode_culprit_pair(int(c)))
            out.append(pairs)
        return out

    def fk(self, q, stream=None, verdict_path: bool = False) -> torch.Tensor:
        """(n, 11, 12): per link position xyz then row-major rotation.  verdict_path=True evaluates the kinematics the
        way the verdict-bit kernels do internally (hardware sin/cos), to measure their distance from the pose kernel."""
        A, B, q9, n = self._planes(q)
        out = torch.empty((n, 11, 12), dtype=torch.float32, device=self.device)
        fn = self.lib.pv_fk_verdict_path if verdict_path else self.lib.pv_fk
        self._ck(fn(self._h, A.data_ptr(), B.data_ptr(), q9.data_ptr() if q9 is not None else None, n,
                    out.data_ptr(), self._stream(stream)), "pv_fk")
        return out

    def check_edges(self, qa, qb, n_steps: int = 0, resolution: float = pm.VALIDITY_RESOLUTION,
                    out: Optional[torch.Tensor] = None, stream=None) -> torch.Tensor:
        aA, aB, a9, n = self._planes(qa)
        bA, bB, b9, nb = self._planes(qb)
        if n != nb:
            raise PandaValidityError("qa and qb must hold the same number of configurations")
        if out is None:
            out = torch.empty((n + 31) // 32, dtype=torch.int32, device=self.device)
        self._ck(self.lib.pv_check_edges(self._h, aA.data_ptr(), aB.data_ptr(), a9.data_ptr() if a9 is not None else None,
                                         bA.data_ptr(), bB.data_ptr(), b9.data_ptr() if b9 is not None else None, n,
                                         int(n_steps), float(resolution), out.data_ptr(), self._stream(stream)),
                 "pv_check_edges")
        return out

    def edge_margins(self, qa, qb, n_steps: int = 0, resolution: float = pm.VALIDITY_RESOLUTION, stream=None):
        aA, aB, a9, n = self._planes(qa)
        bA, bB, b9, nb = self._planes(qb)
        if n != nb:
            raise PandaValidityError("qa and qb must hold the same number of configurations")
        m = torch.empty(n, dtype=torch.float32, device=self.device)
        self._ck(self.lib.pv_edge_margins(self._h, aA.data_ptr(), aB.data_ptr(), a9.data_ptr() if a9 is not None else None,
                                          bA.data_ptr(), bB.data_ptr(), b9.data_ptr() if b9 is not None else None, n,
                                          int(n_steps), float(resolution), m.data_ptr(), self._stream(stream)),
                 "pv_edge_margins")
        return m

    def sweep(self, first: int, n: int, seed: int, fingers_open: bool = True, want_configs: bool = False, stream=None):
        words = (n + 31) // 32
        bits = torch.empty(words, dtype=torch.int32, device=self.device)
        count = torch.zeros(1, dtype=torch.int64, device=self.device)
        qo = torch.empty((n, 9), dtype=torch.float32, device=self.device) if want_configs else None
        self._ck(self.lib.pv_sweep(self._h, int(first), int(n), int(seed) & 0xFFFFFFFF, 1 if fingers_open else 0,
                                   bits.data_ptr(), count.data_ptr(), qo.data_ptr() if qo is not None else None,
                                   self._stream(stream)), "pv_sweep")
        return (bits, count, qo) if want_configs else (bits, count)

    # -- host-buffer calls (what a reference-side binding uses) ----------------------------------------------
    def check_states_host(self, q: np.ndarray, out: Optional[np.ndarray] = None) -> np.ndarray:
        """q: (n, 9) float32 host array (numpy, or a pinned torch tensor's .numpy()).  Returns uint32 words."""
        q = np.ascontiguousarray(q, dtype=np.float32).reshape(-1, 9)
        n = q.shape[0]
        if out is None:
            out = np.empty((n + 31) // 32, dtype=np.uint32)
        self._ck(self.lib.pv_check_states_host(self._h, q.ctypes.data, n, out.ctypes.data), "pv_check_states_host")
        return out

    def check_states_host_arm(self, q7: np.ndarray, fingers=(0.04, 0.04), out: Optional[np.ndarray] = None) -> np.ndarray:
        """q7: (n, 7) float32 host rows of arm joint values; `fingers` = (q8, q9) shared by all of them.  Same verdict
        words as check_states_host on the 9-column rows, 28 B per configuration over PCIe instead of 36."""
        q7 = np.ascontiguousarray(q7, dtype=np.float32).reshape(-1, 7)
        n = q7.shape[0]
        if out is None:
            out = np.empty((n + 31) // 32, dtype=np.uint32)
        self._ck(self.lib.pv_check_states_host_arm(self._h, q7.ctypes.data, n, float(np.float32(fingers[0])),
                                                   float(np.float32(fingers[1])), out.ctypes.data),
                 "pv_check_states_host_arm")
        return out

    def check_edges_host(self, qa: np.ndarray, qb: np.ndarray, n_steps: int = 0,
                         resolution: float = pm.VALIDITY_RESOLUTION, out: Optional[np.ndarray] = None) -> np.ndarray:
        qa = np.ascontiguousarray(qa, dtype=np.float32).reshape(-1, 9)
        qb = np.ascontiguousarray(qb, dtype=np.float32).reshape(-1, 9)
        n = qa.shape[0]
        if qb.shape[0] != n:
            raise PandaValidityError("qa and qb must hold the same number of configurations")
        if out is None:
            out = np.empty((n + 31) // 32, dtype=np.uint32)
        self._ck(self.lib.pv_check_edges_host(self._h, qa.ctypes.data, qb.ctypes.data, n, int(n_steps), float(resolution),
                                              out.ctypes.data), "pv_check_edges_host")
        return out

    def is_state_valid(self, q: Sequence[float]) -> bool:
        """One state through the host entry point (the shape of planning.py:209 `_is_ompl_state_valid`)."""
        w = self.check_states_host(np.asarray(q, dtype=np.float32).reshape(1, 9))
        return bool(w[0] & 1)

    def rrtc_batch(self, starts: np.ndarray, goals: np.ndarray, max_iters: int = 2000, max_nodes: int = 2048,
                   max_path: int = 128, seed: int = 1, replicas: int = 1, shortcut_passes: int = 2,
                   rrt_range: float = 0.0, resolution: float = 0.0, check_endpoints: bool = False,
                   planner: str = "RRTConnect", query_offset: int = 0, packed: bool = False):
        """Batched multi-query planning.  `query_offset` = id of starts[0] in the caller's whole batch: the random
        streams are keyed by global query id, so shards of a batch (other calls, other GPUs) return exactly the rows
        the unsplit call returns -- for any `replicas` (the winner of a query is the search with the smallest
        (iterations, replica id), not the first to finish).
        Returns (paths (n, max_path, 9), lengths, iters, checks); packed=True returns (states (sum(lengths), 9),
        offsets, lengths, iters, checks) instead -- the form that scales to 10^6 queries."""
        starts = np.ascontiguousarray(starts, dtype=np.float32).reshape(-1, 9)
        goals = np.ascontiguousarray(goals, dtype=np.float32).reshape(-1, 9)
        nq = starts.shape[0]
        if goals.shape[0] != nq:
            raise PandaValidityError("starts and goals must have the same length")
        prm = _cabi.PvRrtcParams(float(rrt_range), float(resolution), int(max_iters), int(max_nodes), int(max_path),
                                 int(seed) & 0xFFFFFFFF, int(replicas), int(shortcut_passes), 1 if check_endpoints else 0,
                                 {"RRTConnect": 0, "RRT": 1}[planner], int(query_offset))
        plen = np.zeros(nq, dtype=np.int32)
        iters = np.zeros(nq, dtype=np.int32)
        checks = np.zeros(nq, dtype=np.int64)
        if packed:
            off = np.zeros(nq, dtype=np.int64)
            total = C.c_longlong(0)
            cap = max(4 * nq, 64)  # typical paths hold 2..4 states; grow once if this batch needs more
            while True:
                states = np.empty((cap, 9), dtype=np.float32)
                rc = self.lib.pv_rrtc_batch_packed(self._h, starts.ctypes.data, goals.ctypes.data, nq, C.byref(prm),
                                                   states.ctypes.data, cap, off.ctypes.data, plen.ctypes.data,
                                                   iters.ctypes.data, checks.ctypes.data, C.byref(total))
                if rc == _cabi.PV_ERR_CAPACITY and total.value > cap:
                    cap = int(total.value)
                    continue
                self._ck(rc, "pv_rrtc_batch_packed")
                return states[: total.value], off, plen, iters, checks
        if max_path < 2:
            raise PandaValidityError("max_path must be at least 2")
        paths = np.zeros((nq, max_path, 9), dtype=np.float32)
        self._ck(self.lib.pv_rrtc_batch(self._h, starts.ctypes.data, goals.ctypes.data, nq, C.byref(prm),
                                        paths.ctypes.data, plen.ctypes.data, iters.ctypes.data, checks.ctypes.data),
                 "pv_rrtc_batch")
        return paths, plen, iters, checks

    # -- device steps of the sharded-tree planner front end (distributed.ShardedTreePlanner) ---------------------
    def nn_candidates(self, trees: torch.Tensor, sizes: torch.Tensor, tree_of: Optional[torch.Tensor],
                      targets: torch.Tensor, rank: int, world: int, stream=None) -> torch.Tensor:
        """Nearest node of THIS rank's slots of tree tree_of[t] to targets[t]: (n, 11) = squared distance, global node
        index (int bits), node state.  trees (T, 9, capacity) float32, sizes (T,) int32 slots in use."""
        n = int(targets.shape[0])
        out = torch.empty((n, 11), dtype=torch.float32, device=self.device)
        self._ck(self.lib.pv_nn_candidates(self._h, trees.data_ptr(), sizes.data_ptr(),
                                           tree_of.data_ptr() if tree_of is not None else None, targets.data_ptr(), n,
                                           int(trees.shape[2]), int(rank), int(world), out.data_ptr(),
                                           self._stream(stream)), "pv_nn_candidates")
        return out

    def nn_candidates_gather(self, trees: torch.Tensor, sizes: torch.Tensor, tree_of: Optional[torch.Tensor],
                             targets: torch.Tensor, rank: int, world: int, peer_ptrs_dev: int, multicast_ptr: int,
                             n_peers: Optional[int] = None, stream=None):
        """nn_candidates whose records land in every rank's symmetric buffer at [rank][pair][11] (pv_nn_candidates_gather)."""
        n = int(targets.shape[0])
        self._ck(self.lib.pv_nn_candidates_gather(self._h, trees.data_ptr(), sizes.data_ptr(),
                                                  tree_of.data_ptr() if tree_of is not None else None, targets.data_ptr(), n,
                                                  int(trees.shape[2]), int(rank), int(world),
                                                  C.c_void_p(peer_ptrs_dev or None), int(world if n_peers is None else n_peers),
                                                  C.c_void_p(multicast_ptr or None), self._stream(stream)),
                 "pv_nn_candidates_gather")

    def rrtc_steer(self, cand: torch.Tensor, targets: torch.Tensor, rrt_range: float = 0.0, stream=None):
        """cand (world, n, 11): the all-gathered candidate records.  Returns (from_gidx (n,) int32, ea (n, 9), eb (n, 9),
        reach (n,) int32): the motion og.RRTConnect would validate next (planning.py:156)."""
        world, n = int(cand.shape[0]), int(cand.shape[1])
        gi = torch.empty(n, dtype=torch.int32, device=self.device)
        ea = torch.empty((n, 9), dtype=torch.float32, device=self.device)
        eb = torch.empty((n, 9), dtype=torch.float32, device=self.device)
        reach = torch.empty(n, dtype=torch.int32, device=self.device)
        self._ck(self.lib.pv_rrtc_steer(self._h, cand.data_ptr(), world, n, targets.data_ptr(), float(rrt_range),
                                        gi.data_ptr(), ea.data_ptr(), eb.data_ptr(), reach.data_ptr(),
                                        self._stream(stream)), "pv_rrtc_steer")
        return gi, ea, eb, reach

    def rrtc_samples(self, seed: int, gsearch: torch.Tensor, it: torch.Tensor, stream=None) -> torch.Tensor:
        """Sample it[i] of global search gsearch[i] (int32 tensors) of the planner's random stream: (n, 9)."""
        n = int(gsearch.shape[0])
        out = torch.empty((n, 9), dtype=torch.float32, device=self.device)
        self._ck(self.lib.pv_rrtc_samples(self._h, int(seed) & 0xFFFFFFFF, gsearch.data_ptr(), it.data_ptr(), n,
                                          out.data_ptr(), self._stream(stream)), "pv_rrtc_samples")
        return out

    def plan_path(self, start, goal, num_waypoints: int = 100, smooth: bool = True, planner: str = "RRTConnect",
                  seed: int = 1, replicas: int = 32, max_iters: int = 2000, max_nodes: int = 2048, validate: bool = True,
                  max_attempts: int = 4, timeout: float = 5.0, rrt_range: float = 0.0, resolution: float = 0.0):
        """The whole of planning.py:59-207 in one C call (pv_plan_path): intake check, solve, simplifySolution,
        interpolate, dense validation with fallback / re-planning.  start, goal: 9 joint values (kept in fp64).
        Returns (waypoints (n, 9) float32 -- n = 0 when no path was found --, stats dict)."""
        qs = np.ascontiguousarray(start, dtype=np.float64).reshape(9)
        qg = np.ascontiguousarray(goal, dtype=np.float64).reshape(9)
        # the parameter block is kept between calls (this wrapper is inside every plan's wall time): only what changed
        # since the last call is rewritten
        key = (rrt_range, resolution, max_iters, max_nodes, replicas, smooth, planner, validate, max_attempts, timeout)
        if key != self._plan_key:
            self._plan_prm = _cabi.PvPlanParams(float(rrt_range), float(resolution), int(max_iters), int(max_nodes), 0,
                                                int(replicas), 1 if smooth else 0, {"RRTConnect": 0, "RRT": 1}[planner],
                                                1 if validate else 0, int(max_attempts), float(timeout))
            self._plan_key = key
        prm = self._plan_prm
        prm.seed = int(seed) & 0xFFFFFFFF
        cap = max(int(num_waypoints) if num_waypoints else 0, 256)
        out = np.empty((cap, 9), dtype=np.float32)
        n = C.c_int(0)
        st = _cabi.PvPlanStats()
        self._ck(self.lib.pv_plan_path(self._h, qs.ctypes.data, qg.ctypes.data, int(num_waypoints or 0), C.byref(prm),
                                       out.ctypes.data, cap, C.byref(n), C.byref(st)), "pv_plan_path")
        return out[: n.value], st.as_dict()

    def simplify_path(self, path: np.ndarray, seed: int = 1, resolution: float = 0.0, capacity: int = 256) -> np.ndarray:
        """ss.simplifySolution() (planning.py:196) on a vertex list (n, 9): the simplifier of pv_plan_path on its own."""
        pts = np.ascontiguousarray(np.asarray(path, dtype=np.float64).reshape(-1, 9))
        out = np.empty((capacity, 9), dtype=np.float64)
        n = C.c_int(0)
        counters = (C.c_int * 4)()
        self._ck(self.lib.pv_simplify_path(self._h, pts.ctypes.data, pts.shape[0], int(seed) & 0xFFFFFFFF,
                                           float(resolution), out.ctypes.data, capacity, C.byref(n), counters),
                 "pv_simplify_path")
        self.last_simplify_counters = dict(zip(("partial_rounds", "bspline_steps", "reduce_rounds", "motions"),
                                               [int(c) for c in counters]))
        return out[: n.value].copy()

    def ik_batch(self, pos: np.ndarray, quat: np.ndarray, q_init: Sequence[float], n_seeds: int = 128,
                 max_iters: int = 64, pos_tol: float = 1e-4, rot_tol: float = 1e-3, seed: int = 1):
        """Collision-aware IK for the hand: pos (n, 3), quat (n, 4) wxyz -> (q (n, 9), ok (n,), err (n, 2))."""
        pos = np.ascontiguousarray(pos, dtype=np.float32).reshape(-1, 3)
        quat = np.ascontiguousarray(quat, dtype=np.float32).reshape(-1, 4)
        n = pos.shape[0]
        if quat.shape[0] != n:
            raise PandaValidityError("pos and quat must have the same length")
        qi = np.ascontiguousarray(q_init, dtype=np.float32).reshape(9)
        q = np.zeros((n, 9), dtype=np.float32)
        ok = np.zeros(n, dtype=np.int32)
        err = np.zeros((n, 2), dtype=np.float32)
        self._ck(self.lib.pv_ik_batch(self._h, pos.ctypes.data, quat.ctypes.data, n, qi.ctypes.data, int(n_seeds),
                                      int(max_iters), float(pos_tol), float(rot_tol), int(seed) & 0xFFFFFFFF,
                                      q.ctypes.data, ok.ctypes.data, err.ctypes.data), "pv_ik_batch")
        return q, ok.astype(bool), err

    def ik(self, pos, quat, q_init, **kw) -> Optional[np.ndarray]:
        """One pose, the shape of `robot.inverse_kinematics(link=hand, pos=, quat=)`: qpos (9,) or None."""
        q, ok, _ = self.ik_batch(np.asarray(pos)[None], np.asarray(quat)[None], q_init, **kw)
        return q[0].astype(np.float64) if ok[0] else None

    def fp32_peak(self, iters: int = 4096) -> Tuple[float, float]:
        tf, ms = C.c_double(0), C.c_float(0)
        self._ck(self.lib.pv_fp32_peak(self._h, int(iters), C.byref(tf), C.byref(ms)), "pv_fp32_peak")
        return tf.value, ms.value


def unpack_bits(words, n: int) -> np.ndarray:
    """uint32 verdict words -> bool[n] (bit i&31 of word i>>5)."""
    if torch.is_tensor(words):
        words = words.detach().cpu().numpy()
    w = np.asarray(words).view(np.uint32)
    b = (w[:, None] >> np.arange(32, dtype=np.uint32)) & np.uint32(1)
    return b.reshape(-1)[:n].astype(bool)


def _link_name(i: int) -> str:
    """Link ids 0..10 are the Panda bodies; 11 is the box carried by the hand (pv_set_carried)."""
    return pm.LINK_NAMES[i] if i < len(pm.LINK_NAMES) else "carried_object"


def decode_culprit_pair(code: int):
    """(link name, other name) of a culprit code; other is a link name, 'ground', 'box<k>' or 'joint_limit'."""
    kind, a, b = (code >> 16) & 0xFF, (code >> 8) & 0xFF, code & 0xFF
    if kind == 1:
        return (_link_name(a), "ground")
    if kind == 2:
        return (_link_name(a), f"box{b}")
    if kind == 3:
        return (_link_name(a), _link_name(b))
    if kind == 4:
        return (f"joint{a + 1}", "joint_limit")
    return ("none", "none")


def decode_culprit(code: int) -> str:
    kind, a, b = (code >> 16) & 0xFF, (code >> 8) & 0xFF, code & 0xFF
    if kind == 1:
        return f"{_link_name(a)} vs ground plane"
    if kind == 2:
        return f"{_link_name(a)} vs scene box {b}"
    if kind == 3:
        return f"{_link_name(a)} vs {_link_name(b)}"
    if kind == 4:
        return f"joint {a + 1} outside its limits"
    return "none"
