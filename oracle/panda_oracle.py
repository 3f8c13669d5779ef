"""CPU ORACLE (test infrastructure, NOT product code) -- numpy fp64 restatement of the reference's
state-/motion-validity path for the Franka Panda.

    PARITY UNPINNED vs the reference's Genesis/OMPL verdicts: the reference has no tests, no golden
    vectors, and Genesis / OMPL / the Panda meshes are not installable here.
    What IS pinned: the forward kinematics against the analytic known-answer values of SURVEY.md
    App. A (public Franka kinematics), and the qualitative acceptance facts of App. F (start poses
    and grasp poses the reference plans from/to are valid).  Collision verdicts are pinned only
    between this file, oracle/panda_oracle.c and the CUDA kernels (same primitive model).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import
this module.  The product package never does.

What is restated, and from where:
  * verdict rule                    planning.py:209-219 (_is_ompl_state_valid)
  * attached-object forgiveness     planning.py:221-230 (collision_with_attached_object)
  * joint-limit rule                planning.py:32-41,139-150,165-173 (bounds from robot.q_limit)
  * FK chain constants              scenes.py:85 -> Menagerie panda.xml (un-vendored)
  * base lift                       scenes.py:29-34
  * motion validity                 planning.py:151-156 -> OMPL DiscreteMotionValidator defaults
                                    (un-vendored, un-pinned `ompl` wheel, README.md:53; App. D)
  * path resampling                 planning.py:198 -> OMPL PathGeometric::interpolate(count)
  * Genesis pair semantics          SURVEY App. C (robot-vs-plane, robot-vs-box, robot-vs-robot pairs;
                                    link0-vs-plane never tested because both are fixed)

The collision *primitives* (which spheres/boxes) are an input (`model` dict of arrays, see
rbe550_final_project_b200.panda_model.model_arrays); the chain constants are restated here.
"""
from __future__ import annotations

import numpy as np

# --- kinematic chain (Menagerie panda.xml as loaded at scenes.py:85) -----------------
_POS = np.array([
    [0, 0, 0], [0, 0, 0.333], [0, 0, 0], [0, -0.316, 0], [0.0825, 0, 0], [-0.0825, 0.384, 0],
    [0, 0, 0], [0.088, 0, 0], [0, 0, 0.107], [0, 0, 0.0584], [0, 0, 0.0584],
], dtype=np.float64)
_QUAT = np.array([
    [1, 0, 0, 0], [1, 0, 0, 0], [1, -1, 0, 0], [1, 1, 0, 0], [1, 1, 0, 0], [1, -1, 0, 0],
    [1, 1, 0, 0], [1, 1, 0, 0], [0.9238795, 0, 0, -0.3826834], [1, 0, 0, 0], [0, 0, 0, 1],
], dtype=np.float64)
_PARENT = [-1, 0, 1, 2, 3, 4, 5, 6, 7, 8, 8]
# joint type per body: 'r' revolute about local z with q[j], 's' slide along local +y with q[j]
_JOINT = [None, ("r", 0), ("r", 1), ("r", 2), ("r", 3), ("r", 4), ("r", 5), ("r", 6), None, ("s", 7), ("s", 8)]

SAT_PARALLEL_EPS = 1e-4  # cross axes with |a_i x b_j|^2 below this are skipped (edges ~parallel)


def _quat_mat(q):
    w, x, y, z = q / np.linalg.norm(q)
    return np.array([
        [1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
        [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
        [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)],
    ])


_ROT = np.stack([_quat_mat(q) for q in _QUAT])


def fk(q, base=(0.0, 0.0, 0.01)):
    """Forward kinematics of the 11 Panda bodies.

    q: (n, 9).  Returns R (n, 11, 3, 3) world-from-link rotations and p (n, 11, 3) origins.
    child = parent * Trans(pos) * Rot(quat) * [Rot_z(q_j) | Trans_y(q_j)]   (MuJoCo body/joint rule).
    """
    q = np.atleast_2d(np.asarray(q, dtype=np.float64))
    n = q.shape[0]
    R = np.zeros((n, 11, 3, 3))
    p = np.zeros((n, 11, 3))
    for i in range(11):
        if _PARENT[i] < 0:
            Rw = np.broadcast_to(_ROT[i], (n, 3, 3)).copy()
            pw = np.broadcast_to(np.asarray(base, dtype=np.float64) + _POS[i], (n, 3)).copy()
        else:
            Rp, pp = R[:, _PARENT[i]], p[:, _PARENT[i]]
            Rw = Rp @ _ROT[i]
            pw = pp + Rp @ _POS[i]
        jt = _JOINT[i]
        if jt is not None:
            kind, j = jt
            if kind == "r":
                c, s = np.cos(q[:, j]), np.sin(q[:, j])
                Rz = np.zeros((n, 3, 3))
                Rz[:, 0, 0], Rz[:, 0, 1], Rz[:, 1, 0], Rz[:, 1, 1], Rz[:, 2, 2] = c, -s, s, c, 1.0
                Rw = Rw @ Rz
            else:
                pw = pw + Rw[:, :, 1] * q[:, j:j + 1]
        R[:, i], p[:, i] = Rw, pw
    return R, p


# --- primitive distances (signed margins; < 0 means penetration) ---------------------------------------
HAND_LINK = 8
CARRY_LAST_ARM_LINK = 6  # the carried box is tested against the spheres of link0..link6


def _sphere_obb_margin(c, r, bc, bh, bR):
    """c (n,3) sphere centres, r scalar or (n,), box centre bc (.., 3), half bh (..,3), rot bR (..,3,3)
    (world-from-box).  Returns margin (n,)."""
    d = c - bc
    loc = np.einsum("...ji,...j->...i", bR, d)  # R^T d
    e = np.abs(loc) - bh
    out = np.maximum(e, 0.0)
    dist = np.sqrt((out * out).sum(-1))
    inside = e.max(-1)  # negative when the centre is inside the box
    return np.where(dist > 0.0, dist, inside) - r


def _obb_obb_margin(ca, ha, Ra, cb, hb, Rb):
    """SAT margin between boxes A (n-batched) and B (single).  Largest normalised separation over the
    15 axes; cross axes with near-parallel edges are skipped.  < 0 on overlap."""
    Rm = np.einsum("nji,jk->nik", Ra, Rb)  # Ra^T Rb
    t = np.einsum("nji,nj->ni", Ra, cb - ca)
    A = np.abs(Rm)
    best = np.full(ca.shape[0], -np.inf)
    for i in range(3):  # A's face normals
        ra = ha[i]
        rb = (A[:, i, :] * hb).sum(-1)
        best = np.maximum(best, np.abs(t[:, i]) - ra - rb)
    for j in range(3):  # B's face normals
        ra = (A[:, :, j] * ha).sum(-1)
        rb = hb[j]
        tl = (t * Rm[:, :, j]).sum(-1)
        best = np.maximum(best, np.abs(tl) - ra - rb)
    for i in range(3):
        i1, i2 = (i + 1) % 3, (i + 2) % 3
        for j in range(3):
            j1, j2 = (j + 1) % 3, (j + 2) % 3
            len2 = 1.0 - Rm[:, i, j] ** 2
            ra = ha[i1] * A[:, i2, j] + ha[i2] * A[:, i1, j]
            rb = hb[j1] * A[:, i, j2] + hb[j2] * A[:, i, j1]
            tl = np.abs(t[:, i2] * Rm[:, i1, j] - t[:, i1] * Rm[:, i2, j])
            ok = len2 > SAT_PARALLEL_EPS
            sep = np.where(ok, (tl - ra - rb) / np.sqrt(np.where(ok, len2, 1.0)), -np.inf)
            best = np.maximum(best, sep)
    return best


def state_margin(q, scene, model, attached=-1, self_collision=True, base=(0.0, 0.0, 0.01), detail=False):
    """Minimum signed clearance (m) of each configuration; the state is valid iff margin >= 0.

    scene: dict(obb=(B,16) rows [c.xyz, half.xyz, R row-major 9, pad], table_z=float)
    attached: scene-box index whose contacts with hand / fingers are forgiven (planning.py:221-230);
              -1 for none.  The attached box stays a static obstacle for every other link (App. E-3).
    scene["carried"] (optional, NOT reference behaviour: SURVEY.md 8f-3 / App. E-3 "physically-correct mode"):
              dict(index=k, R=(3,3), t=(3,), shrink=float) -- scene box k is rigidly held by the hand with pose
              hand-from-box (R, t); in its own tests its half extents are reduced by `shrink` (contact allowance:
              a block resting on the table or on another block is not in collision with it).  It then stops being a static obstacle and instead moves with the hand: it is tested
              against the plane, every other scene box and the arm spheres of link0..link6 (link7, hand and
              fingers are rigid or in grasp contact with it).
    Follows planning.py:209-219: any robot contact invalidates the state unless forgiven.
    """
    q = np.atleast_2d(np.asarray(q, dtype=np.float64))
    n = q.shape[0]
    R, p = fk(q, base)
    obb = np.asarray(scene["obb"], dtype=np.float64).reshape(-1, 16)
    tz = float(scene["table_z"])
    sl, sc, sr = model["sphere_link"], model["sphere_center"], model["sphere_radius"]
    bl, bcn, bh = model["box_link"], model["box_center"], model["box_half"]
    S, H = len(sl), len(bl)
    margin = np.full(n, np.inf)
    parts = {}

    def take(name, m):
        nonlocal margin
        margin = np.minimum(margin, m)
        if detail:
            parts[name] = np.minimum(parts.get(name, np.inf), m)

    wc = np.stack([p[:, sl[i]] + R[:, sl[i]] @ sc[i] for i in range(S)], axis=1)  # (n,S,3)
    bw = np.stack([p[:, bl[k]] + R[:, bl[k]] @ bcn[k] for k in range(H)], axis=1)  # (n,H,3)
    bR = np.stack([R[:, bl[k]] for k in range(H)], axis=1)  # (n,H,3,3)

    # robot vs ground plane (link0 is fixed to the world -> pair filtered, App. C / App. F-5)
    for i in range(S):
        if sl[i] != 0:
            take("table", wc[:, i, 2] - sr[i] - tz)
    for k in range(H):
        ext = (np.abs(bR[:, k, 2, :]) * bh[k]).sum(-1)
        take("table", bw[:, k, 2] - ext - tz)

    carried = scene.get("carried") if isinstance(scene, dict) else None
    ck = -1
    if carried is not None:
        ck = int(carried["index"])
        ch = obb[ck, 3:6] - float(carried.get("shrink", 0.0))
        cR = R[:, HAND_LINK] @ np.asarray(carried["R"], dtype=np.float64).reshape(3, 3)  # (n,3,3) world-from-box
        cc = p[:, HAND_LINK] + R[:, HAND_LINK] @ np.asarray(carried["t"], dtype=np.float64)
        take("carried", cc[:, 2] - (np.abs(cR[:, 2, :]) * ch).sum(-1) - tz)
        for i in range(S):
            if sl[i] <= CARRY_LAST_ARM_LINK:
                take("carried", _sphere_obb_margin(wc[:, i], sr[i], cc, ch, cR))

    # robot vs scene boxes
    for b in range(obb.shape[0]):
        if b == ck:
            continue  # carried: it is where the hand is, not where the snapshot saw it
        cb, hb, Rb = obb[b, 0:3], obb[b, 3:6], obb[b, 6:15].reshape(3, 3)
        if ck >= 0:
            take("carried", _obb_obb_margin(cc, ch, cR, cb, hb, Rb))
        for i in range(S):
            take("env", _sphere_obb_margin(wc[:, i], sr[i], cb, hb, Rb))
        if b == attached:
            continue
        for k in range(H):
            take("env", _obb_obb_margin(bw[:, k], bh[k], bR[:, k], cb, hb, Rb))

    if self_collision:
        for a, b in model["ss_pairs"]:
            d = np.sqrt(((wc[:, a] - wc[:, b]) ** 2).sum(-1))
            take("self", d - (sr[a] + sr[b]))
        for a, k in model["sb_pairs"]:
            take("self", _sphere_obb_margin(wc[:, a], sr[a], bw[:, k], bh[k], bR[:, k]))
    # joint limits (planning.py:139-150) are part of the validity domain: OMPL never hands the callback a state outside
    # them, and the pruned self-pair lists are certified inside them only.  Compared in fp32 (the value the kernels,
    # whose inputs are fp32, see), so a state AT a limit is inside in either precision; the negated comparison also
    # rejects non-finite joint values
    lo32, hi32 = model["q_lower"].astype(np.float32), model["q_upper"].astype(np.float32)
    with np.errstate(invalid="ignore", over="ignore"):
        q32 = q.astype(np.float32)
        outside = ~((q32 >= lo32) & (q32 <= hi32)).all(axis=1)
    margin = np.where(outside, -1e30, margin)
    if detail:
        return margin, parts
    return margin


def in_bounds(q, model):
    """OMPL RealVectorStateSpace::satisfiesBounds with the bounds of planning.py:139-150."""
    q = np.atleast_2d(np.asarray(q, dtype=np.float64))
    eps = np.finfo(np.float64).eps
    return np.all((q - eps <= model["q_upper"]) & (q + eps >= model["q_lower"]), axis=1)


def state_valid(q, scene, model, attached=-1, self_collision=True, check_limits=False):
    m = state_margin(q, scene, model, attached, self_collision)
    v = m >= 0.0
    if check_limits:
        v &= in_bounds(q, model)
    return v


# --- motion validity -----------------------------------------------------------------------------------
def edge_steps(qa, qb, n_steps=0, resolution=0.13037159046356686):
    """Number of states checked on an edge.  n_steps > 0: fixed (BASELINE config 3: t=(k+1)/n).
    n_steps == 0: OMPL DiscreteMotionValidator rule nd = ceil(|a-b| / resolution) (App. D);
    states checked are t = k/nd, k = 1..nd (endpoint b included, a assumed valid)."""
    if n_steps > 0:
        return np.full(np.atleast_2d(qa).shape[0], n_steps, dtype=np.int64)
    d = np.sqrt(((np.atleast_2d(qb).astype(np.float64) - np.atleast_2d(qa)) ** 2).sum(-1))
    return np.maximum(np.ceil(d / resolution).astype(np.int64), 1)


def edge_margin(qa, qb, scene, model, n_steps=0, attached=-1, self_collision=True,
                resolution=0.13037159046356686):
    """Min margin over the states q(t) = qa + t (qb - qa), t = k/nd, k=1..nd.  Edge valid iff >= 0.
    Order of evaluation (OMPL checks b first then bisects) does not change the AND."""
    qa = np.atleast_2d(np.asarray(qa, dtype=np.float64))
    qb = np.atleast_2d(np.asarray(qb, dtype=np.float64))
    nd = edge_steps(qa, qb, n_steps, resolution)
    out = np.full(qa.shape[0], np.inf)
    for k in range(1, int(nd.max()) + 1):
        sel = np.nonzero(nd >= k)[0]
        t = (k / nd[sel])[:, None]
        qs = qa[sel] + t * (qb[sel] - qa[sel])
        out[sel] = np.minimum(out[sel], state_margin(qs, scene, model, attached, self_collision))
    return out


# --- OMPL PathGeometric::interpolate(count) (planning.py:198) ----------------------------------------
def interpolate_path(states, count):
    """Restatement of ompl::geometric::PathGeometric::interpolate(unsigned count) for a
    RealVectorStateSpace (linear interpolation, L2 distance).  Returns an (m, d) array, m >= count
    unless the path already has more than `count` states (then unchanged)."""
    st = [np.asarray(s, dtype=np.float64) for s in states]
    if count < len(st) or len(st) < 2:
        return np.array(st)
    count = int(count)
    seg = [float(np.linalg.norm(st[i + 1] - st[i])) for i in range(len(st) - 1)]
    remaining = float(sum(seg))
    new = []
    n1 = len(st) - 1
    for i in range(n1):
        s1, s2 = st[i], st[i + 1]
        new.append(s1)
        max_n = count + i - len(st)
        if max_n > 0:
            if i + 1 == n1:
                ns = max_n + 2
            else:
                ns = int(np.floor(0.5 + count * seg[i] / remaining)) + 1 if remaining > 0 else 2
            if ns > 2:
                ns -= 2
                if ns > max_n:
                    ns = max_n
                for j in range(1, ns + 1):
                    new.append(s1 + (j / (ns + 1)) * (s2 - s1))
            else:
                ns = 0
            count -= ns + 1
            remaining -= seg[i]
        else:
            count -= 1
    new.append(st[n1])
    return np.array(new)


# --- counter-based RNG for the device-generated sweeps (BASELINE config 5) ------------------------------
_PH_M0, _PH_M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
_PH_W0, _PH_W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)


def philox4x32(counter, key):
    """Philox-4x32-10 (Salmon et al., SC'11).  counter (n,4) uint32, key (2,) uint32 -> (n,4) uint32."""
    c = np.array(counter, dtype=np.uint32).reshape(-1, 4).copy()
    k0, k1 = np.uint32(key[0]), np.uint32(key[1])
    mask = np.uint64(0xFFFFFFFF)
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = _PH_M0 * c[:, 0].astype(np.uint64)
            p1 = _PH_M1 * c[:, 2].astype(np.uint64)
            hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), (p0 & mask).astype(np.uint32)
            hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), (p1 & mask).astype(np.uint32)
            c = np.stack([hi1 ^ c[:, 1] ^ k0, lo1, hi0 ^ c[:, 3] ^ k1, lo0], axis=1)
            k0 = np.uint32(k0 + _PH_W0)
            k1 = np.uint32(k1 + _PH_W1)
    return c


def sweep_configs(first, n, seed, model, fingers_open=True):
    """Config i of the sweep (fp32): three Philox blocks keyed by (seed, 0x50414e44) with counter
    (i_lo, i_hi, blk, 0); u = (x >> 8) * 2^-24; q_j = fma(u_j, hi_j - lo_j, lo_j) in fp32.
    fingers_open: q8 = q9 = 0.04 (BASELINE config 2), else sampled like the arm joints."""
    idx = np.arange(first, first + n, dtype=np.uint64)
    lo32 = (idx & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    hi32 = (idx >> np.uint64(32)).astype(np.uint32)
    key = (np.uint32(seed & 0xFFFFFFFF), np.uint32(0x50414E44))
    u = []
    for blk in range(3):
        ctr = np.stack([lo32, hi32, np.full(n, blk, np.uint32), np.zeros(n, np.uint32)], axis=1)
        r = philox4x32(ctr, key)
        u.append((r >> np.uint32(8)).astype(np.float32) * np.float32(2.0 ** -24))
    u = np.concatenate(u, axis=1)[:, :9]
    lo = model["q_lower"].astype(np.float32)
    hi = model["q_upper"].astype(np.float32)
    span = (hi - lo).astype(np.float32)
    # fma in fp32 == round(fp64 exact product + addend) because u*span is exact in fp64 (24+24 bits)
    q = (u.astype(np.float64) * span.astype(np.float64) + lo.astype(np.float64)).astype(np.float32)
    if fingers_open:
        q[:, 7] = np.float32(0.04)
        q[:, 8] = np.float32(0.04)
    return q


def pack_bits(valid):
    """Verdict bit layout of the C-ABI: bit (i & 31) of word (i >> 5), 1 = valid."""
    v = np.asarray(valid, dtype=bool)
    n = v.shape[0]
    pad = (-n) % 32
    v = np.concatenate([v, np.zeros(pad, bool)])
    w = v.reshape(-1, 32).astype(np.uint32) << np.arange(32, dtype=np.uint32)
    return w.sum(axis=1, dtype=np.uint64).astype(np.uint32)


def unpack_bits(words, n):
    w = np.asarray(words, dtype=np.uint32)
    b = (w[:, None] >> np.arange(32, dtype=np.uint32)) & np.uint32(1)
    return b.reshape(-1)[:n].astype(bool)
