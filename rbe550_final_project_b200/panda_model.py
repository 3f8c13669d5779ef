"""Frozen Franka Panda model: kinematic chain, joint limits, collision primitives, self-pair list.

This is *data*, not a code path.  Everything the CUDA kernels unroll at compile time
(`csrc/panda_model_gen.h`) is generated from the tables below by `write_header()`, and the
tests hand the very same tables to the CPU oracle, so kernel and oracle always see one model.

Provenance
----------
* Kinematic tree and joint limits: the MuJoCo-Menagerie `panda.xml` that the reference loads at
  scenes.py:85 (`gs.morphs.MJCF(file="xml/franka_emika_panda/panda.xml")`); values transcribed in
  SURVEY.md App. A.  The file itself is not in /root/reference nor in this image.
* Base lift of +0.01 m: scenes.py:29-34 (`_elevate_robot_base`).
* Link names "hand" / "left_finger" / "right_finger": planning.py:222, motion_primitives.py:132.
* Collision primitives: AUTHORED HERE from the public link dimensions.  The
  reference collides Genesis' convex hulls of the Menagerie meshes, which are not available, so the
  sphere chains (capsules sampled as swept spheres) and the three gripper boxes below are an
  approximation of that geometry.  They are frozen: S, H, P below define the algorithmic FLOP count
  used by bench.py's roofline.

Link indices: 0..7 = link0..link7, 8 = hand, 9 = left_finger, 10 = right_finger.
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass
from typing import List, Tuple

import numpy as np

N_Q = 9  # 7 revolute + 2 prismatic finger joints (planning.py:143, scenes.py:92)
N_LINKS = 11
LINK_NAMES = (
    "link0", "link1", "link2", "link3", "link4", "link5", "link6", "link7",
    "hand", "left_finger", "right_finger",
)
LINK_HAND, LINK_LF, LINK_RF = 8, 9, 10
PARENT = (-1, 0, 1, 2, 3, 4, 5, 6, 7, 8, 8)

# joint limits, SURVEY.md App. A (Menagerie panda.xml `range` attributes)
Q_LOWER = np.array([-2.8973, -1.7628, -2.8973, -3.0718, -2.8973, -0.0175, -2.8973, 0.0, 0.0])
Q_UPPER = np.array([2.8973, 1.7628, 2.8973, -0.0698, 2.8973, 3.7525, 2.8973, 0.04, 0.04])

# OMPL defaults the reference relies on (planning.py:151-156 -> SimpleSetup defaults)
SPACE_EXTENT = float(np.linalg.norm(Q_UPPER - Q_LOWER))  # 13.03716
VALIDITY_RESOLUTION = 0.01 * SPACE_EXTENT  # DiscreteMotionValidator: 1 % of extent
RRTC_RANGE = 0.2 * SPACE_EXTENT  # RRTConnect default range: 20 % of extent

# body placement in the parent frame: (pos, quat wxyz un-normalised as written in the MJCF)
BODY_POS = (
    (0.0, 0.0, 0.0),
    (0.0, 0.0, 0.333),
    (0.0, 0.0, 0.0),
    (0.0, -0.316, 0.0),
    (0.0825, 0.0, 0.0),
    (-0.0825, 0.384, 0.0),
    (0.0, 0.0, 0.0),
    (0.088, 0.0, 0.0),
    (0.0, 0.0, 0.107),
    (0.0, 0.0, 0.0584),
    (0.0, 0.0, 0.0584),
)
BODY_QUAT = (
    (1.0, 0.0, 0.0, 0.0),
    (1.0, 0.0, 0.0, 0.0),
    (1.0, -1.0, 0.0, 0.0),
    (1.0, 1.0, 0.0, 0.0),
    (1.0, 1.0, 0.0, 0.0),
    (1.0, -1.0, 0.0, 0.0),
    (1.0, 1.0, 0.0, 0.0),
    (1.0, 1.0, 0.0, 0.0),
    (0.9238795, 0.0, 0.0, -0.3826834),
    (1.0, 0.0, 0.0, 0.0),
    (0.0, 0.0, 0.0, 1.0),
)
BASE_LIFT = (0.0, 0.0, 0.01)  # scenes.py:29-34

# Named configurations used by the reference scripts
Q_SAFE_HOME = np.array([0.0, -0.785, 0.0, -2.356, 0.0, 1.571, 0.785, 0.04, 0.04])  # goal1_scattered.py:43
Q_SAFE_HOME_039 = np.array([0.0, -0.785, 0.0, -2.356, 0.0, 1.571, 0.785, 0.039, 0.039])  # goal4_task1.py:40
Q_SCENE_INIT = np.array([0.0, -0.5, -0.2, -1.0, 0.0, 1.0, 0.5, 0.02, 0.02])  # scenes.py:92


# ---------------------------------------------------------------------------------------------
# Collision primitives
# ---------------------------------------------------------------------------------------------
@dataclass(frozen=True)
class Capsule:
    """A capsule in a link frame, sampled into `n` equal spheres (end points included)."""
    link: int
    p0: Tuple[float, float, float]
    p1: Tuple[float, float, float]
    r: float
    n: int


# Derivation (SURVEY.md App. F envelopes): each arm link is one or two capsules around the joint
# axes / the span to the next joint; `n` is chosen so neighbouring samples are <= ~1 radius apart.
ARM_CAPSULES: Tuple[Capsule, ...] = (
    # link0: base block, ~0.2 m long behind/under joint 1, top at z ~ 0.14
    Capsule(0, (-0.09, 0.0, 0.06), (0.0, 0.0, 0.06), 0.085, 2),
    # link1: vertical shoulder column below the joint-1 frame + joint-2 housing on -y
    Capsule(1, (0.0, 0.0, -0.18), (0.0, 0.0, 0.0), 0.06, 4),
    Capsule(1, (0.0, -0.06, 0.0), (0.0, -0.06, 0.0), 0.06, 1),
    # link2: joint-2 housing on +z, lower half of the upper arm along -y
    Capsule(2, (0.0, 0.0, 0.06), (0.0, 0.0, 0.06), 0.06, 1),
    Capsule(2, (0.0, 0.0, 0.0), (0.0, -0.18, 0.0), 0.06, 4),
    # link3: upper half of the upper arm along -z, elbow housing at x = 0.0825 (+y side)
    Capsule(3, (0.0, 0.0, -0.12), (0.0, 0.0, 0.0), 0.06, 3),
    Capsule(3, (0.0825, 0.045, 0.0), (0.0825, 0.045, 0.0), 0.055, 1),
    # link4: elbow housing along +z, body towards the forearm at (-0.0825, ~0.1, 0)
    Capsule(4, (0.0, 0.0, 0.0), (0.0, 0.0, 0.06), 0.055, 2),
    Capsule(4, (-0.04, 0.05, 0.0), (-0.0825, 0.10, 0.0), 0.057, 2),
    # link5: forearm (frame at the wrist end, z along the forearm), thin offset bar, wrist housing
    Capsule(5, (0.0, 0.0, -0.26), (0.0, 0.0, -0.20), 0.06, 2),
    Capsule(5, (0.0, 0.04, -0.15), (0.0, 0.08, -0.05), 0.045, 3),
    Capsule(5, (0.0, 0.0, 0.0), (0.0, 0.06, 0.0), 0.055, 2),
    # link6: joint-6 housing, bridge to joint 7 at x = 0.088, joint-7 housing on +y
    Capsule(6, (0.0, 0.0, 0.02), (0.0, 0.0, 0.02), 0.055, 1),
    Capsule(6, (0.044, 0.0, 0.0), (0.088, 0.0, 0.0), 0.055, 2),
    Capsule(6, (0.088, 0.045, 0.0), (0.088, 0.045, 0.0), 0.05, 1),
    # link7: flange body between joint 7 and the hand mount at z = 0.107
    Capsule(7, (0.0, 0.0, 0.03), (0.0, 0.0, 0.08), 0.045, 2),
)


def _sample_spheres() -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    link, ctr, rad = [], [], []
    for cap in ARM_CAPSULES:
        p0, p1 = np.array(cap.p0), np.array(cap.p1)
        for k in range(cap.n):
            t = 0.0 if cap.n == 1 else k / (cap.n - 1)
            link.append(cap.link)
            ctr.append(p0 + t * (p1 - p0))
            rad.append(cap.r)
    return np.array(link, dtype=np.int32), np.array(ctr, dtype=np.float64), np.array(rad, dtype=np.float64)


SPHERE_LINK, SPHERE_CENTER, SPHERE_RADIUS = _sample_spheres()
N_SPHERES = int(SPHERE_LINK.shape[0])

# Gripper boxes: (link, centre in the link frame, half extents).  Hand ~0.063 x 0.204 x 0.066 from the
# hand origin to past the finger mount; each finger ~0.021 x 0.018 x 0.054 whose inner (pad) face is
# the plane y = 0 of its own frame, so the gap between the pads is q8 + q9.
BOX_LINK = np.array([LINK_HAND, LINK_LF, LINK_RF], dtype=np.int32)
BOX_CENTER = np.array([(0.0, 0.0, 0.033), (0.0, 0.009, 0.027), (0.0, 0.009, 0.027)], dtype=np.float64)
BOX_HALF = np.array([(0.0315, 0.102, 0.033), (0.0105, 0.009, 0.027), (0.0105, 0.009, 0.027)], dtype=np.float64)
N_BOXES = 3

# Links whose contacts with the attached object are forgiven (planning.py:222)
ATTACH_FORGIVEN_LINKS = (LINK_HAND, LINK_LF, LINK_RF)


# ---------------------------------------------------------------------------------------------
# fp64 helpers used only to *derive* the frozen pair list (not the oracle, not the product path)
# ---------------------------------------------------------------------------------------------
def _quat_to_mat(q) -> np.ndarray:
    w, x, y, z = np.asarray(q, dtype=np.float64) / np.linalg.norm(q)
    return np.array([
        [1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
        [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
        [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)],
    ])


def _neutral_frames() -> List[Tuple[np.ndarray, np.ndarray]]:
    """Link frames at the neutral pose q = 0 (used for the neutral-overlap pair filter only)."""
    frames = []
    for i in range(N_LINKS):
        Rl, pl = _quat_to_mat(BODY_QUAT[i]), np.array(BODY_POS[i])
        if PARENT[i] < 0:
            frames.append((Rl, pl))
        else:
            Rp, pp = frames[PARENT[i]]
            frames.append((Rp @ Rl, pp + Rp @ pl))
    return frames


def _sphere_box_dist(c, bc, bh, bR) -> float:
    loc = bR.T @ (c - bc)
    d = np.maximum(np.abs(loc) - bh, 0.0)
    return float(np.linalg.norm(d))


def _adjacent(la: int, lb: int) -> bool:
    return PARENT[la] == lb or PARENT[lb] == la


def _derive_pairs():
    """Static self-collision pair filter: skip same-link pairs, skip parent-child
    link pairs, skip pairs already overlapping at the neutral pose q = 0."""
    fr = _neutral_frames()
    wc = np.array([fr[l][1] + fr[l][0] @ c for l, c in zip(SPHERE_LINK, SPHERE_CENTER)])
    ss = []
    for a in range(N_SPHERES):
        for b in range(a + 1, N_SPHERES):
            la, lb = int(SPHERE_LINK[a]), int(SPHERE_LINK[b])
            if la == lb or _adjacent(la, lb):
                continue
            if np.linalg.norm(wc[a] - wc[b]) < SPHERE_RADIUS[a] + SPHERE_RADIUS[b]:
                continue
            ss.append((a, b))
    sb = []
    for k in range(N_BOXES):
        lk = int(BOX_LINK[k])
        Rk, pk = fr[lk]
        bc = pk + Rk @ BOX_CENTER[k]
        for a in range(N_SPHERES):
            la = int(SPHERE_LINK[a])
            if la == lk or _adjacent(la, lk):
                continue
            if _sphere_box_dist(wc[a], bc, BOX_HALF[k], Rk) < SPHERE_RADIUS[a]:
                continue
            sb.append((a, k))
    # box-box: hand-finger pairs are parent-child; left-right finger pads touch at the neutral pose
    # (q8 = q9 = 0 -> inner faces coincide at y = 0), so that pair is removed by the neutral filter.
    return np.array(ss, dtype=np.int32), np.array(sb, dtype=np.int32)


def derive_pairs_unpruned():
    """Pair lists after the static App. C filter only (before the certified never-collide pruning)."""
    return _derive_pairs()


def model_fingerprint() -> str:
    """Hash of everything the never-collide certificate depends on."""
    import hashlib
    h = hashlib.sha256()
    for arr in (SPHERE_LINK, SPHERE_CENTER, SPHERE_RADIUS, BOX_LINK, BOX_CENTER, BOX_HALF, Q_LOWER, Q_UPPER,
                np.array(BODY_POS), np.array(BODY_QUAT)):
        h.update(np.ascontiguousarray(arr).tobytes())
    return h.hexdigest()[:16]


NEVER_COLLIDE_PATH = os.path.join(os.path.dirname(__file__), "data", "never_collide.json")


def _prune_certified(ss, sb):
    """Drop the pairs that tools/certify_never_collide.py proved can never touch within the joint limits
    (branch-and-bound with Lipschitz bounds, certified clearance >= 2 mm).  The certificate is tied to the
    primitive tables by a fingerprint; a stale or missing file means no pruning."""
    import json
    try:
        with open(NEVER_COLLIDE_PATH) as fh:
            cert = json.load(fh)
    except FileNotFoundError:
        return ss, sb, None
    if cert.get("model_fingerprint") != model_fingerprint():
        return ss, sb, None
    never_ss = {tuple(p) for p in cert["never_ss"]}
    never_sb = {tuple(p) for p in cert["never_sb"]}
    ss2 = np.array([p for p in ss if (int(p[0]), int(p[1])) not in never_ss], dtype=np.int32).reshape(-1, 2)
    sb2 = np.array([p for p in sb if (int(p[0]), int(p[1])) not in never_sb], dtype=np.int32).reshape(-1, 2)
    return ss2, sb2, cert


SS_PAIRS_UNPRUNED, SB_PAIRS_UNPRUNED = _derive_pairs()
SS_PAIRS, SB_PAIRS, NEVER_COLLIDE_CERT = _prune_certified(SS_PAIRS_UNPRUNED, SB_PAIRS_UNPRUNED)
N_SS_PAIRS = int(SS_PAIRS.shape[0])
N_SB_PAIRS = int(SB_PAIRS.shape[0])


# ---------------------------------------------------------------------------------------------
# Algorithmic FLOPs per state check, frozen with this model
# ---------------------------------------------------------------------------------------------
F_FK = 460
F_PLACE = 18
F_SPHERE_BOX = 30
F_PLANE = 2
F_SPHERE_SPHERE = 9
F_BOX_BOX = 200


def flops_per_state_check(n_obb: int) -> int:
    """Full no-early-exit work of one valid configuration against `n_obb` scene boxes."""
    s_env = N_SPHERES * (F_PLACE + n_obb * F_SPHERE_BOX + F_PLANE)
    b_env = N_BOXES * (F_PLACE + n_obb * F_BOX_BOX + 8 * F_PLANE)
    self_ = N_SS_PAIRS * F_SPHERE_SPHERE + N_SB_PAIRS * F_SPHERE_BOX
    return F_FK + s_env + b_env + self_


# ---------------------------------------------------------------------------------------------
# Header generator for the CUDA side
# ---------------------------------------------------------------------------------------------
def _f(x: float) -> str:
    s = repr(float(np.float32(x)))
    if "e" not in s and "." not in s:
        s += ".0"
    return s + "f"


CULL_SLACK = 1e-4  # metres added to every culling radius so fp32 rounding can never cull a true contact


def bounding_ball(idx):
    """(centre sphere, radius) of the smallest ball centred on one of the spheres `idx` (all on one link) that contains
    all of them.  Distances inside a link do not depend on the configuration."""
    best = None
    for c in idx:
        rad = max(float(np.linalg.norm(SPHERE_CENTER[i] - SPHERE_CENTER[c]) + SPHERE_RADIUS[i]) for i in idx)
        if best is None or rad < best[1]:
            best = (c, rad)
    return best


def link_groups():
    """Per arm link: (link, centre sphere index, bounding radius) of a ball, centred on one of the
    link's own sphere centres, that contains all spheres of that link.  Used for conservative culling."""
    out = []
    for l in range(8):
        idx = [i for i in range(N_SPHERES) if int(SPHERE_LINK[i]) == l]
        if not idx:
            continue
        best = None
        for c in idx:
            rad = max(float(np.linalg.norm(SPHERE_CENTER[i] - SPHERE_CENTER[c]) + SPHERE_RADIUS[i]) for i in idx)
            if best is None or rad < best[1]:
                best = (c, rad)
        out.append((l, best[0], best[1]))
    return out


BOX_BOUND_RADIUS = np.linalg.norm(BOX_HALF, axis=1)
FINGER_SLIDE_MAX = float(Q_UPPER[7])


def static_reach():
    """Conservative reach of each link group (links 1..7) and gripper box (hand, left, right) measured from the
    shoulder point S0 = base + (0, 0, 0.333) (origin of link1 and link2): no point of the group can ever be
    farther from S0, whatever the joint angles (triangle inequality over the body offsets).  pv_set_scene uses
    it to drop (group, scene box) combinations that cannot touch -- a uniform, divergence-free cull.
    link0 does not move at all and is tested exactly on the host."""
    chain = np.zeros(N_LINKS)
    for l in range(3, N_LINKS):
        chain[l] = chain[PARENT[l]] + float(np.linalg.norm(BODY_POS[l]))
    chain[LINK_LF] += FINGER_SLIDE_MAX
    chain[LINK_RF] += FINGER_SLIDE_MAX
    link_reach = np.zeros(8)
    for l in range(1, 8):
        idx = [i for i in range(N_SPHERES) if int(SPHERE_LINK[i]) == l]
        link_reach[l] = chain[l] + max(float(np.linalg.norm(SPHERE_CENTER[i]) + SPHERE_RADIUS[i]) for i in idx)
    box_reach = np.array([chain[int(BOX_LINK[k])] + float(np.linalg.norm(BOX_CENTER[k]) + BOX_BOUND_RADIUS[k])
                          for k in range(N_BOXES)])
    return link_reach, box_reach


def motion_reach_bounds():
    """R_j for the seven revolute joints: no point the validity tests look at (arm sphere centres, gripper box centres
    and corners) is ever farther than R_j from joint j's axis, whatever the configuration (triangle inequality over the
    body offsets from the joint's own origin, which lies on its axis, plus the finger travel).  When the joints move by
    dq, every such point therefore moves by at most sum_j |dq_j| R_j + |dq_8| + |dq_9| -- in the world and, for a point
    distal to link a, relative to link a's frame (joints up to a move both rigidly).  The motion validator turns that
    into a certificate: a state whose culling tests clear by more than this displacement proves its neighbours on the
    motion valid without testing them (pv_edge_kernel, csrc/pv_edge.cu)."""
    loc = np.zeros(N_LINKS)
    for i in range(N_SPHERES):
        l = int(SPHERE_LINK[i])
        loc[l] = max(loc[l], float(np.linalg.norm(SPHERE_CENTER[i])))
    for k in range(N_BOXES):
        l = int(BOX_LINK[k])
        loc[l] = max(loc[l], float(np.linalg.norm(np.abs(BOX_CENTER[k]) + BOX_HALF[k])))
    out = np.zeros(7)
    for j in range(1, 8):  # joint j turns body j about the z axis through body j's origin
        for b in range(j, N_LINKS):
            chain, l = 0.0, b
            while l != j and l >= 0:
                chain += float(np.linalg.norm(BODY_POS[l])) + (FINGER_SLIDE_MAX if l in (LINK_LF, LINK_RF) else 0.0)
                l = PARENT[l]
            if l == j:
                out[j - 1] = max(out[j - 1], chain + loc[b])
    return out


MOTION_CERT_MAX_SLACK = 0.10  # metres: motions whose displacement bound between tested states exceeds this seek no certificate


def _place_expr(c, axis):
    """p.axis + X.axis*cx + Y.axis*cy + Z.axis*cz with zero terms dropped (fmaf chain)."""
    e = f"p.{axis}"
    for vec, v in (("X", c[0]), ("Y", c[1]), ("Z", c[2])):
        if float(np.float32(v)) != 0.0:
            e = f"fmaf({vec}.{axis}, {_f(v)}, {e})"
    return e


def header_text() -> str:
    L = []
    a = L.append
    a("// GENERATED by rbe550_final_project_b200/panda_model.py (write_header) -- do not edit.")
    a("// Frozen Panda collision model; see panda_model.py for provenance.")
    a("#pragma once")
    a(f"#define PV_N_Q {N_Q}")
    a(f"#define PV_N_LINKS {N_LINKS}")
    a(f"#define PV_N_SPHERES {N_SPHERES}")
    a(f"#define PV_N_BOXES {N_BOXES}")
    a(f"#define PV_N_SS_PAIRS {N_SS_PAIRS}")
    a(f"#define PV_N_SB_PAIRS {N_SB_PAIRS}")
    a(f"#define PV_SPACE_EXTENT {_f(SPACE_EXTENT)}")
    a(f"#define PV_VALIDITY_RESOLUTION {_f(VALIDITY_RESOLUTION)}")
    a(f"#define PV_RRTC_RANGE {_f(RRTC_RANGE)}")
    a(f"#define PV_CULL_SLACK {_f(CULL_SLACK)}")
    hq = np.array(BODY_QUAT[LINK_HAND]) / np.linalg.norm(BODY_QUAT[LINK_HAND])
    ang = 2.0 * math.atan2(hq[3], hq[0])
    a(f"#define PV_HAND_COS {_f(math.cos(ang))}")
    a(f"#define PV_HAND_SIN {_f(math.sin(ang))}")
    lr, br_ = static_reach()
    a("// conservative reach from the shoulder point (base + (0,0,0.333)) of link groups 0..7 (entry 0 unused:")
    a("// link0 is fixed and tested exactly) and of the 3 gripper boxes")
    a("#define PV_LINK_REACH {" + ", ".join(_f(v) for v in lr) + "}")
    a("#define PV_BOX_REACH {" + ", ".join(_f(v) for v in br_) + "}")
    a("#define PV_SPHERE_LINK {" + ", ".join(str(int(v)) for v in SPHERE_LINK) + "}")
    a("// motion certificates (motion_reach_bounds): R_j of the seven revolute joints, and the largest slack ever sought")
    a("#define PV_MOTION_REACH {" + ", ".join(_f(v * 1.0001) for v in motion_reach_bounds()) + "}")
    a(f"#define PV_MOTION_CERT_MAX_SLACK {_f(MOTION_CERT_MAX_SLACK)}")
    a("// the largest radius (sum) any self-collision test compares a distance with: sphere pairs, sphere vs gripper box")
    rs_ = max([float(SPHERE_RADIUS[a_] + SPHERE_RADIUS[b_]) for a_, b_ in SS_PAIRS] + [float(SPHERE_RADIUS[a_]) for a_, _ in SB_PAIRS])
    a(f"#define PV_SELF_R_MAX {_f(rs_ * 1.0001)}")
    a("// joint limits")
    a("#define PV_Q_LOWER {" + ", ".join(_f(v) for v in Q_LOWER) + "}")
    a("#define PV_Q_UPPER {" + ", ".join(_f(v) for v in Q_UPPER) + "}")
    a("// X(idx, link, cx, cy, cz, r): arm spheres in their link frame")
    a("#define PV_SPHERES(X) \\")
    for i in range(N_SPHERES):
        c = SPHERE_CENTER[i]
        a(f"  X({i}, {int(SPHERE_LINK[i])}, {_f(c[0])}, {_f(c[1])}, {_f(c[2])}, {_f(SPHERE_RADIUS[i])}) \\")
    a("")
    for l in range(8):
        a(f"#define PV_SPHERES_LINK{l}(X) \\")
        for i in range(N_SPHERES):
            if int(SPHERE_LINK[i]) == l:
                c = SPHERE_CENTER[i]
                a(f"  X({i}, {l}, {_f(c[0])}, {_f(c[1])}, {_f(c[2])}, {_f(SPHERE_RADIUS[i])}) \\")
        a("")
        a(f"// world centres of link{l}'s spheres from its frame (p, X, Y, Z); zero terms dropped")
        a(f"#define PV_PLACE_LINK{l}(s, p, X, Y, Z) \\")
        for i in range(N_SPHERES):
            if int(SPHERE_LINK[i]) == l:
                c = SPHERE_CENTER[i]
                a(f"  s[{i}] = make_float3({_place_expr(c, 'x')}, {_place_expr(c, 'y')}, {_place_expr(c, 'z')}); \\")
        a("")
    a("// lowest point of the moving arm spheres (link1..link7) for the ground-plane test: min over spheres of z - r,")
    a("// evaluated as one min-chain per distinct radius (rounding is monotonic, so the verdict is bit-identical to")
    a("// testing every sphere on its own)")
    by_r = {}
    for i in range(N_SPHERES):
        if int(SPHERE_LINK[i]) != 0:
            by_r.setdefault(_f(SPHERE_RADIUS[i]), []).append(i)
    def min_tree(items):  # balanced, so the chain of dependent FMNMX stays short
        if len(items) == 1:
            return items[0]
        h = (len(items) + 1) // 2
        return f"fminf({min_tree(items[:h])}, {min_tree(items[h:])})"
    terms = [f"({min_tree([f's[{i}].z' for i in idx])} - {r})" for r, idx in by_r.items()]
    a(f"#define PV_TABLE_LOWEST(s) ({min_tree(terms)})")
    a("// X(link, centre_sphere, bound_radius): conservative bounding ball of each link's spheres")
    a("#define PV_LINK_GROUPS(X) \\")
    groups = link_groups()
    for l, c, r in groups:
        a(f"  X({l}, {c}, {_f(r)}) \\")
    a("")
    a("// X(k, link, cx, cy, cz, hx, hy, hz, bound_radius): gripper boxes in their link frame")
    a("#define PV_BOXES(X) \\")
    for k in range(N_BOXES):
        c, h = BOX_CENTER[k], BOX_HALF[k]
        a(f"  X({k}, {int(BOX_LINK[k])}, {_f(c[0])}, {_f(c[1])}, {_f(c[2])}, {_f(h[0])}, {_f(h[1])}, {_f(h[2])}, {_f(BOX_BOUND_RADIUS[k])}) \\")
    a("")
    g = {l: (c, r) for l, c, r in groups}
    a("// sphere-sphere self pairs grouped by link pair.  LP(la, lb, ca, cb, cull_r2): the block of pairs")
    a("// PV_SS_PAIRS_la_lb can be skipped when |s[ca]-s[cb]|^2 >= cull_r2.  X(a, b, (ra+rb)^2, ra+rb)")
    a("#define PV_SS_LINKPAIRS(LP) \\")
    lps = sorted({(int(SPHERE_LINK[p]), int(SPHERE_LINK[q])) for p, q in SS_PAIRS})
    for la, lb in lps:
        # balls around the spheres that actually take part in this block (often a few at one end of the link), not
        # around the whole links: link2-vs-link5 passes for 1.7 % of random configurations instead of 33 %
        ca, ra = bounding_ball(sorted({int(p) for p, q in SS_PAIRS if (int(SPHERE_LINK[p]), int(SPHERE_LINK[q])) == (la, lb)}))
        cb, rb = bounding_ball(sorted({int(q) for p, q in SS_PAIRS if (int(SPHERE_LINK[p]), int(SPHERE_LINK[q])) == (la, lb)}))
        rr = ra + rb + CULL_SLACK
        a(f"  LP({la}, {lb}, {ca}, {cb}, {_f(rr * rr)}) \\")
    a("")
    for la, lb in lps:
        a(f"#define PV_SS_PAIRS_{la}_{lb}(X) \\")
        for p, q in SS_PAIRS:
            if (int(SPHERE_LINK[p]), int(SPHERE_LINK[q])) == (la, lb):
                rs = float(np.float32(SPHERE_RADIUS[p] + SPHERE_RADIUS[q]))
                a(f"  X({p}, {q}, {_f(rs * rs)}, {_f(rs)}) \\")
        a("")
    # ---- the same pairs, two per instruction (sm_100a packed FP32: FADD2 / FFMA2) -----------------------------
    a("// Packed form of the pair lists above: X2(a, b0, b1, n0, n1, k) tests sphere a against spheres b0 and b1 in one go")
    a("// (k = running index mod 4, for callers that spread the results over independent accumulators);")
    a("// n = -(ra+rb)^2, or 0 for a half that is not on the list (then d^2 + 0 < 0 never fires).  b0, b1 are neighbours")
    a("// inside one link (b1 == b0 for the odd one out), so each sphere has ONE register partner.")
    for la, lb in lps:
        pairs = {(int(p), int(q)) for p, q in SS_PAIRS if (int(SPHERE_LINK[p]), int(SPHERE_LINK[q])) == (la, lb)}
        best = None
        for swap in (False, True):  # broadcast side = la (pack lb) or lb (pack la)
            pl = lb if not swap else la
            idx = [i for i in range(N_SPHERES) if int(SPHERE_LINK[i]) == pl]
            slots = [(idx[k], idx[k + 1] if k + 1 < len(idx) else idx[k]) for k in range(0, len(idx), 2)]
            bl = la if not swap else lb
            rows = []
            for s_ in [i for i in range(N_SPHERES) if int(SPHERE_LINK[i]) == bl]:
                for b0, b1 in slots:
                    def rr2(x):
                        key = (s_, x) if not swap else (x, s_)
                        if key not in pairs:
                            return None
                        rs = float(np.float32(SPHERE_RADIUS[key[0]] + SPHERE_RADIUS[key[1]]))
                        return rs * rs
                    n0, n1 = rr2(b0), (rr2(b1) if b1 != b0 else None)
                    if n0 is None and n1 is None:
                        continue
                    rows.append((s_, b0, b1, -n0 if n0 is not None else 0.0, -n1 if n1 is not None else 0.0))
            if best is None or len(rows) < len(best):
                best = rows
        a(f"#define PV_SS2_PAIRS_{la}_{lb}(X2) \\")
        for k_, (s_, b0, b1, n0, n1) in enumerate(best):
            a(f"  X2({s_}, {b0}, {b1}, {_f(n0)}, {_f(n1)}, {k_ % 4}) \\")
        a("")
    # ---- sphere-vs-gripper self pairs evaluated in the hand frame ---------------------------------------------
    a("// gripper boxes in the HAND frame: X(k, cx, cy0, cz, slide_sign, finger_q_index, hx, hy, hz); the centre's y is")
    a("// cy0 + slide_sign * q[finger_q_index] (hand: sign 0).  All three boxes share the hand's axes.")
    a("#define PV_BOXES_HANDFRAME(X) \\")
    for k in range(N_BOXES):
        c, h = BOX_CENTER[k], BOX_HALF[k]
        if k == 0:
            a(f"  X(0, {_f(c[0])}, {_f(c[1])}, {_f(c[2])}, 0.0f, 7, {_f(h[0])}, {_f(h[1])}, {_f(h[2])}) \\")
        else:
            sign = 1.0 if k == 1 else -1.0  # right finger frame = hand frame turned half a turn about z
            off = np.array(BODY_POS[int(BOX_LINK[k])])
            a(f"  X({k}, {_f(sign * c[0] + off[0])}, {_f(sign * c[1] + off[1])}, {_f(c[2] + off[2])}, {_f(sign)}, {6 + k}, "
              f"{_f(h[0])}, {_f(h[1])}, {_f(h[2])}) \\")
    a("")
    # while both fingers are within +-FINGER_SLIDE_MAX of the hand's mid-plane the hand box's own bounding ball already
    # contains both finger boxes, so the per-configuration gripper ball of pv_check_config is a constant there
    for k in (1, 2):
        off = np.array(BODY_POS[int(BOX_LINK[k])])
        for qf in (-FINGER_SLIDE_MAX, FINGER_SLIDE_MAX):
            c = np.array([BOX_CENTER[k][0] + off[0], BOX_CENTER[k][1] + qf + off[1], BOX_CENTER[k][2] + off[2]])
            assert np.linalg.norm(c - BOX_CENTER[0]) + BOX_BOUND_RADIUS[k] < BOX_BOUND_RADIUS[0] - 1e-4
    a("// |q7|, |q8| <= this: the gripper ball (centre = hand box centre) is the hand box's own bounding ball")
    a(f"#define PV_GRIP_CONST_MAXQ {_f(FINGER_SLIDE_MAX)}")
    a("// LB(la, ca, cull0, cull1, cull2, r_la): link la's block runs when its bounding ball (centre s[ca], radius r_la)")
    a("// reaches the bounding ball of gripper box k for any k (cull_k = squared reach, 0 when the link has no pair with box k)")
    a("#define PV_SBH_LINKS(LB) \\")
    sb_links = sorted({int(SPHERE_LINK[p]) for p, _ in SB_PAIRS})
    for la in sb_links:
        culls = []
        ca, ra = bounding_ball(sorted({int(p) for p, _ in SB_PAIRS if int(SPHERE_LINK[p]) == la}))
        for k in range(N_BOXES):
            if any(int(SPHERE_LINK[p]) == la and int(kk) == k for p, kk in SB_PAIRS):
                rr = ra + BOX_BOUND_RADIUS[k] + CULL_SLACK
                culls.append(_f(rr * rr))
            else:
                culls.append("0.0f")
        a(f"  LB({la}, {ca}, {', '.join(culls)}, {_f(ra + CULL_SLACK)}) \\")
    a("")
    a("// per link: S(a, r, r2) brings sphere a into the hand frame, B(a, k) tests it against gripper box k")
    for la in sb_links:
        a(f"#define PV_SBH_{la}(S, B) \\")
        for sp in sorted({int(p) for p, _ in SB_PAIRS if int(SPHERE_LINK[p]) == la}):
            r = float(np.float32(SPHERE_RADIUS[sp]))
            line = f"  S({sp}, {_f(r)}, {_f(r * r)})"
            for p, kk in SB_PAIRS:
                if int(p) == sp:
                    line += f" B({sp}, {int(kk)})"
            a(line + " \\")
        a("")
    return "\n".join(L) + "\n"


HEADER_PATH = os.path.join(os.path.dirname(__file__), "csrc", "panda_model_gen.h")


def write_header(path: str = HEADER_PATH) -> bool:
    """Write the generated header if it changed. Returns True if the file was (re)written."""
    txt = header_text()
    try:
        with open(path) as fh:
            if fh.read() == txt:
                return False
    except FileNotFoundError:
        pass
    with open(path, "w") as fh:
        fh.write(txt)
    return True


def model_arrays(dtype=np.float64) -> dict:
    """The model as plain arrays, in the layout the CPU oracle takes (oracle/panda_oracle.c)."""
    return dict(
        sphere_link=SPHERE_LINK.copy(),
        sphere_center=SPHERE_CENTER.astype(dtype),
        sphere_radius=SPHERE_RADIUS.astype(dtype),
        box_link=BOX_LINK.copy(),
        box_center=BOX_CENTER.astype(dtype),
        box_half=BOX_HALF.astype(dtype),
        ss_pairs=SS_PAIRS.copy(),
        sb_pairs=SB_PAIRS.copy(),
        q_lower=Q_LOWER.astype(dtype),
        q_upper=Q_UPPER.astype(dtype),
    )


if __name__ == "__main__":
    print("spheres", N_SPHERES, "boxes", N_BOXES, "ss pairs", N_SS_PAIRS, "sb pairs", N_SB_PAIRS)
    print("extent", SPACE_EXTENT, "resolution", VALIDITY_RESOLUTION, "range", RRTC_RANGE)
    for b in (6, 10):
        print("flops/check @", b, "obb:", flops_per_state_check(b))
    print("header rewritten:", write_header())
