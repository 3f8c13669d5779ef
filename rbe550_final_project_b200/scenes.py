"""Scene snapshots: the block scenes of the reference as flat OBB buffers for the validity kernels.

A `SceneSnapshot` is what the C-ABI's `pv_set_scene` consumes: up to PV_MAX_OBB oriented boxes
(16 floats each: centre xyz, half extents xyz, world-from-box rotation row-major, bounding radius),
the ground-plane height and the robot base position.

Fixtures reproduce the *numbers* of the reference scene factories (scenes.py:41-373) with the
random xy jitter (scenes.py:36-39, seeded by wall-clock) switched off or replaced by a fixed seed:

  goal1_scattered     scenes.py:52-57            6 blocks, yaw 0                (BASELINE config 1, 2)
  goal1_stacked       scenes.py:109-135          one 6-high column
  goal3_initial       scenes.py:157-166          two rows of five
  goal3_tower         goal3_tallest.py:63-101    8-high tower at (0.45, 0) + r2, o2 on the table (config 4)
  goal4_task1_initial scenes.py:241-252
  goal4_task1_pentagon goal4_task1.py:66-126     finished 2-layer pentagon, true yawed OBBs (config 3)
  goal4_task2         scenes.py:317-322

`snapshot_from_sim(scene, robot)` builds the same buffer from live Genesis-like entities
(duck-typed: `scene.entities`, `entity.morph.size`, `entity.get_pos()`, `entity.get_quat()`), which is
how `PlannerInterface` sees the world when `motion_primitives.py` drives it (planning.py:24-30).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Any, List, Optional, Sequence

import numpy as np

from . import panda_model as pm

PV_MAX_OBB = 32
BLOCK_SIZE = 0.04  # scenes.py:60 Box(size=(0.04, 0.04, 0.04))
BLOCK_Z = 0.02


def quat_wxyz_to_mat(q: Sequence[float]) -> np.ndarray:
    w, x, y, z = np.asarray(q, dtype=np.float64) / np.linalg.norm(q)
    return np.array([
        [1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
        [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
        [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)],
    ])


def yaw_mat(deg: float) -> np.ndarray:
    a = math.radians(deg)
    c, s = math.cos(a), math.sin(a)
    return np.array([[c, -s, 0.0], [s, c, 0.0], [0.0, 0.0, 1.0]])


@dataclass
class SceneSnapshot:
    obb: np.ndarray  # (B, 16) float32
    table_z: float = 0.0
    base: tuple = pm.BASE_LIFT
    names: List[str] = field(default_factory=list)
    entity_idx: List[int] = field(default_factory=list)  # Genesis entity idx of each box (plane = 0)

    @property
    def n_obb(self) -> int:
        return int(self.obb.shape[0])

    def as_oracle_scene(self) -> dict:
        return {"obb": self.obb.astype(np.float64), "table_z": float(self.table_z)}

    def index_of_entity(self, idx: int) -> int:
        """Scene-box slot of the Genesis entity index `idx` (planning.py:226 compares the contact's
        geom index with `attached_object.idx`; plane = 0, block k = k in every reference scene)."""
        try:
            return self.entity_idx.index(int(idx))
        except ValueError:
            return -1


def make_obb(center, size, R=None) -> np.ndarray:
    rec = np.zeros(16, dtype=np.float64)
    half = 0.5 * np.asarray(size, dtype=np.float64)
    rec[0:3] = center
    rec[3:6] = half
    rec[6:15] = (np.eye(3) if R is None else np.asarray(R, dtype=np.float64)).reshape(9)
    rec[15] = float(np.linalg.norm(half))
    return rec


def _from_blocks(blocks, names=None) -> SceneSnapshot:
    """blocks: list of (x, y, z, yaw_deg)."""
    recs = [make_obb((x, y, z), (BLOCK_SIZE,) * 3, yaw_mat(yaw)) for x, y, z, yaw in blocks]
    n = len(recs)
    return SceneSnapshot(
        obb=np.array(recs, dtype=np.float32).reshape(n, 16),
        names=list(names) if names else [f"b{i}" for i in range(n)],
        entity_idx=list(range(1, n + 1)),
    )


def _jitter(xy, rng, noise=0.05):
    if rng is None:
        return xy
    return (xy[0] + rng.uniform(-noise, noise), xy[1] + rng.uniform(-noise, noise))


def goal1_scattered(seed: Optional[int] = None) -> SceneSnapshot:
    rng = None if seed is None else np.random.default_rng(seed)
    nominal = [(0.65, 0.0), (0.65, 0.2), (0.65, 0.4), (0.45, 0.0), (0.45, 0.2), (0.45, 0.4)]
    pts = [_jitter(p, rng) for p in nominal]
    return _from_blocks([(x, y, BLOCK_Z, 0.0) for x, y in pts], names=list("rgbymc"))


def goal1_stacked(seed: Optional[int] = None) -> SceneSnapshot:
    rng = None if seed is None else np.random.default_rng(seed)
    x, y = _jitter((0.45, 0.0), rng, 0.2)
    return _from_blocks([(x, y, 0.02 + 0.04 * k, 0.0) for k in range(6)], names=list("rgbymc"))


def goal3_initial() -> SceneSnapshot:
    names = ["r", "g", "b", "y", "o", "r2", "g2", "b2", "y2", "o2"]
    ys = [-0.4, -0.2, 0.0, 0.2, 0.4]
    blocks = [(0.45, y, BLOCK_Z, 0.0) for y in ys] + [(0.65, y, BLOCK_Z, 0.0) for y in ys]
    return _from_blocks(blocks, names)


def goal3_tower(height: int = 8) -> SceneSnapshot:
    order = ["b", "b2", "g", "y", "g2", "y2", "r", "o", "r2", "o2"]  # goal3_tallest.py:63-80 build order
    blocks = [(0.45, 0.0, 0.02 + 0.04 * k, 0.0) for k in range(height)]
    loose = {"r2": (0.65, -0.4), "o2": (0.65, 0.4), "r": (0.45, -0.4), "o": (0.45, 0.4)}
    names = order[:height]
    for nm in order[height:]:
        x, y = loose[nm]
        blocks.append((x, y, BLOCK_Z, 0.0))
        names.append(nm)
    return _from_blocks(blocks, names)


def goal4_task1_initial() -> SceneSnapshot:
    xy = [(0.35, -0.40), (0.35, -0.25), (0.45, -0.30), (0.6, -0.40), (0.6, -0.25),
          (0.35, 0.40), (0.35, 0.25), (0.45, 0.30), (0.6, 0.40), (0.6, 0.25)]
    return _from_blocks([(x, y, BLOCK_Z, 0.0) for x, y in xy], [f"b{i + 1}" for i in range(10)])


def goal4_task1_pentagon() -> SceneSnapshot:
    """Finished pentagon of goal4_task1.py:66-126: centre (0.50, 0.1), radius 0.06, base layer at
    angles 72 i (x + 0.0045), top layer at 36 + 72 i (x + 0.0053, y - 0.0005), one block higher."""
    cx, cy, rad = 0.50, 0.1, 0.06

    def wrap(a):
        while a < -180:
            a += 360
        while a > 180:
            a -= 360
        return a

    blocks = []
    for i in range(5):
        ang = 72.0 * i
        blocks.append((cx + rad * math.cos(math.radians(ang)) + 0.0045, cy + rad * math.sin(math.radians(ang)),
                       BLOCK_Z, wrap(ang)))
    for i in range(5):
        ang = 72.0 * i + 36.0
        blocks.append((cx + rad * math.cos(math.radians(ang)) + 0.0053,
                       cy + rad * math.sin(math.radians(ang)) - 0.0005, BLOCK_Z + BLOCK_SIZE, wrap(ang)))
    return _from_blocks(blocks, [f"b{i + 1}" for i in range(10)])


def goal4_task2() -> SceneSnapshot:
    xy = [(0.65, 0.0), (0.55, 0.2), (0.6, 0.4), (0.45, 0.0), (0.45, 0.2), (0.45, 0.4)]
    return _from_blocks([(x, y, BLOCK_Z, 0.0) for x, y in xy], ["r1", "r2", "r3", "g1", "g2", "g3"])


FIXTURES = {
    "goal1_scattered": goal1_scattered,
    "goal1_stacked": goal1_stacked,
    "goal3_initial": goal3_initial,
    "goal3_tower": goal3_tower,
    "goal4_task1_initial": goal4_task1_initial,
    "goal4_task1_pentagon": goal4_task1_pentagon,
    "goal4_task2": goal4_task2,
}


# ---------------------------------------------------------------------------------------------
# Snapshot of a live (Genesis-like) scene
# ---------------------------------------------------------------------------------------------
def _to_np(x) -> np.ndarray:
    if hasattr(x, "detach"):
        x = x.detach()
    if hasattr(x, "cpu"):
        x = x.cpu().numpy()
    return np.asarray(x, dtype=np.float64).reshape(-1)


def _host_vec(x):
    """get_pos() / get_quat() of a Genesis-like entity -> something np.array() digests cheaply (torch -> numpy)."""
    if hasattr(x, "detach"):
        x = x.detach()
        if x.device.type != "cpu":
            x = x.cpu()
        return x.numpy()
    return x


_BOX_CACHE: dict = {}
_LAST_SNAPSHOT: dict = {}  # id(scene) -> (box list, positions, quaternions, base, snapshot) of the previous call
_IDENTITY_QUAT = (1.0, 0.0, 0.0, 0.0)


def _box_entities(scene: Any, raw: Any, robot: Any):
    """(entities, names, idxs, half extents) of the box entities of a scene -- everything about them that does not
    change between plans, cached per scene object and entity count."""
    ents = getattr(scene, "entities", [])
    key = (id(scene), len(ents), id(raw))
    hit = _BOX_CACHE.get(key)
    if hit is not None and all(a is b for a, b in zip(hit[0], ents)):
        return hit[1]
    boxes, names, idxs, halves = [], [], [], []
    for k, ent in enumerate(ents):
        if ent is raw or ent is robot:
            continue
        size = getattr(getattr(ent, "morph", None), "size", None)
        if size is None:
            continue
        boxes.append(ent)
        names.append(str(getattr(ent, "name", f"entity{k}")))
        idxs.append(int(getattr(ent, "idx", k)))
        halves.append([0.5 * float(v) for v in size[:3]])
    if len(boxes) > PV_MAX_OBB:
        raise ValueError(f"scene has {len(boxes)} boxes; the validity kernels stage at most {PV_MAX_OBB}")
    half = np.asarray(halves, dtype=np.float64).reshape(len(boxes), 3)
    val = (boxes, names, idxs, half, np.sqrt((half * half).sum(axis=1)), [hasattr(e, "get_quat") for e in boxes])
    if len(_BOX_CACHE) > 64:
        _BOX_CACHE.clear()
    _BOX_CACHE[key] = (list(ents), val)
    return val


def snapshot_from_sim(scene: Any, robot: Any) -> SceneSnapshot:
    """Read every box entity of a Genesis-like scene into an OBB buffer.

    Duck-typed on what Genesis exposes: `scene.entities` (insertion order, plane first,
    scenes.py:49-85), `entity.idx`, `entity.morph.size` for `gs.morphs.Box`, `entity.get_pos()`,
    `entity.get_quat()` (wxyz).  The robot entity (unwrapped through RobotAdapter.raw/.robot), planes
    and anything without a box size are skipped; the plane supplies table_z = 0.
    This runs once per plan_path, i.e. inside the plan time: the per-entity facts that cannot change (which entities
    are boxes, their sizes, names, ids) are cached, the poses are read every time and turned into rotation matrices
    for all boxes at once.
    """
    raw = getattr(robot, "raw", None) or getattr(robot, "robot", robot)
    boxes, names, idxs, half, radius, has_quat = _box_entities(scene, raw, robot)
    n = len(boxes)
    pos = quat = None
    if n:
        pos = np.array([_host_vec(e.get_pos()) for e in boxes], dtype=np.float64)
        quat = np.array([_host_vec(e.get_quat()) if hq else _IDENTITY_QUAT for e, hq in zip(boxes, has_quat)], dtype=np.float64)
        if pos.shape != (n, 3) or quat.shape != (n, 4):
            pos = np.ascontiguousarray(pos.reshape(n, -1)[:, :3])
            quat = np.ascontiguousarray(quat.reshape(n, -1)[:, :4])
    base = pm.BASE_LIFT
    if hasattr(raw, "get_pos"):
        try:
            b = _host_vec(raw.get_pos())
            base = (float(b[0]), float(b[1]), float(b[2]))
        except Exception:
            base = pm.BASE_LIFT
    # the poses are READ on every call (blocks move between plans); when they are what the previous call saw -- plan after
    # plan of one primitive -- the previous snapshot object is handed back and nothing is converted again
    last = _LAST_SNAPSHOT.get(id(scene))
    if last is not None and last[0] is boxes and last[3] == base and (n == 0 or (np.array_equal(last[1], pos) and
                                                                               np.array_equal(last[2], quat))):
        return last[4]
    obb = np.empty((n, 16), dtype=np.float32)
    if n:
        from . import _cabi
        if _cabi.load().pv_obb_from_poses(pos.ctypes.data, quat.ctypes.data, half.ctypes.data, n, obb.ctypes.data) != 0:
            raise ValueError("snapshot_from_sim: an entity reports a zero or non-finite quaternion")
    obb.flags.writeable = False  # a snapshot that may be handed out again must not be edited in place
    snap = SceneSnapshot(obb=obb, table_z=0.0, base=base, names=names, entity_idx=idxs)
    if len(_LAST_SNAPSHOT) > 64:
        _LAST_SNAPSHOT.clear()
    _LAST_SNAPSHOT[id(scene)] = (boxes, pos, quat, base, snap)
    return snap
