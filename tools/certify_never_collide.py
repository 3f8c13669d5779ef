"""Offline, rigorous pruning of self-collision primitive pairs that can NEVER touch within the joint limits.

For a pair (primitive A on link la, primitive B on link lb, la < lb) the clearance depends only on the
joints between the two links.  Branch and bound over that joint box: at a box centre q_c with half-widths
h_j the clearance f obeys   f(q) >= f(q_c) - sum_j L_j h_j   for every q in the box, where L_j bounds how far
any point of B moves per radian of joint j (its largest possible distance from that joint's axis).  A pair
is CERTIFIED never-colliding when the whole joint box is covered by sub-boxes whose lower bound is
>= MARGIN; it is KEPT as soon as one sample has f < MARGIN or the box budget runs out (conservative).

The result (data/never_collide.json) is frozen model data: panda_model.py drops the certified pairs from
the pair lists both the CUDA kernels and the CPU oracle use.  tests/test_model_pruning.py re-checks the
certificate by dense random sampling against the UNPRUNED pair list.

Run:  python tools/certify_never_collide.py        (about a minute on 8 cores)
"""
import json
import os
import sys
import time
from multiprocessing import Pool

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rbe550_final_project_b200 import panda_model as pm  # noqa: E402

MARGIN = 2e-3          # certified clearance floor (m); far above fp32 rounding (1e-6) and the 1e-4 band
MAX_BOXES = 3_000_000  # evaluation budget per pair
ROT = np.stack([pm._quat_to_mat(q) for q in pm.BODY_QUAT])
POS = np.array(pm.BODY_POS)
# finger boxes swept over their slide range, expressed in the hand frame (conservative)
SLIDE = 0.04


def rel_frames(la, lb_arm, Q):
    """Pose of link lb_arm (<= 8) in the frame of link la for joint samples Q (n, lb_arm - la) (hand has no joint)."""
    n = Q.shape[0]
    R = np.broadcast_to(np.eye(3), (n, 3, 3)).copy()
    p = np.zeros((n, 3))
    col = 0
    for k in range(la + 1, lb_arm + 1):
        p = p + R @ POS[k]
        R = R @ ROT[k]
        if 1 <= k <= 7:
            c, s = np.cos(Q[:, col]), np.sin(Q[:, col])
            Rz = np.zeros((n, 3, 3))
            Rz[:, 0, 0], Rz[:, 0, 1], Rz[:, 1, 0], Rz[:, 1, 1], Rz[:, 2, 2] = c, -s, s, c, 1.0
            R = R @ Rz
            col += 1
    return R, p


def joint_cols(la, lb_arm):
    return [k - 1 for k in range(la + 1, min(lb_arm, 7) + 1)]


def lipschitz(la, lb_arm, extent):
    """L_j for each joint between la and lb_arm: max distance of any point of B from the joint axis."""
    out = []
    for k in range(la + 1, min(lb_arm, 7) + 1):
        reach = sum(np.linalg.norm(POS[m]) for m in range(k + 1, lb_arm + 1)) + extent
        out.append(reach)
    return np.array(out)


def box_geom(k):
    """(centre, half) of gripper box k in the HAND frame, fingers swept over the slide range."""
    if k == 0:
        return pm.BOX_CENTER[0].copy(), pm.BOX_HALF[0].copy()
    c, h = pm.BOX_CENTER[k], pm.BOX_HALF[k]
    lo_y, hi_y = c[1] - h[1], c[1] + h[1] + SLIDE
    cy, hy = 0.5 * (lo_y + hi_y), 0.5 * (hi_y - lo_y)
    sign = 1.0 if k == 1 else -1.0  # right finger frame is the hand's turned half a turn about z
    return np.array([0.0, sign * cy, c[2] + POS[9][2]]), np.array([h[0], hy, h[2]])


def clearance_fn(kind, a, b):
    la = int(pm.SPHERE_LINK[a])
    ca, ra = pm.SPHERE_CENTER[a], pm.SPHERE_RADIUS[a]
    if kind == "ss":
        lb = int(pm.SPHERE_LINK[b])
        cb, rb = pm.SPHERE_CENTER[b], pm.SPHERE_RADIUS[b]
        L = lipschitz(la, lb, np.linalg.norm(cb))

        def f(Q):
            R, p = rel_frames(la, lb, Q)
            w = p + R @ cb
            return np.linalg.norm(w - ca, axis=1) - (ra + rb)
        return f, joint_cols(la, lb), L
    bc, bh = box_geom(b)
    L = lipschitz(la, 8, np.linalg.norm(bc) + np.linalg.norm(bh))

    def f(Q):
        R, p = rel_frames(la, 8, Q)
        d = ca - (p + R @ bc)
        loc = np.einsum("nji,nj->ni", R, d)
        e = np.abs(loc) - bh
        out = np.maximum(e, 0.0)
        dist = np.sqrt((out * out).sum(1))
        return np.where(dist > 0, dist, e.max(1)) - ra
    return f, joint_cols(la, 8), L


def certify(task):
    kind, a, b = task
    f, cols, L = clearance_fn(kind, a, b)
    lo, hi = pm.Q_LOWER[cols], pm.Q_UPPER[cols]
    rng = np.random.default_rng(1000 * a + b)
    # 1) cheap refutation by sampling
    Q = rng.uniform(lo, hi, size=(20000, len(cols)))
    fmin = float(f(Q).min())
    if fmin < MARGIN:
        return (kind, a, b, False, fmin, 0)
    # 2) branch and bound
    centres = (0.5 * (lo + hi))[None, :]
    halves = (0.5 * (hi - lo))[None, :]
    used = 0
    while centres.shape[0]:
        used += centres.shape[0]
        if used > MAX_BOXES:
            return (kind, a, b, False, fmin, used)
        val = f(centres)
        fmin = min(fmin, float(val.min()))
        if fmin < MARGIN:
            return (kind, a, b, False, fmin, used)
        bound = val - (halves * L[None, :]).sum(1)
        keep = bound < MARGIN
        centres, halves = centres[keep], halves[keep]
        if not centres.shape[0]:
            break
        dim = np.argmax(halves * L[None, :], axis=1)
        idx = np.arange(centres.shape[0])
        h2 = halves.copy()
        h2[idx, dim] *= 0.5
        c1, c2 = centres.copy(), centres.copy()
        c1[idx, dim] -= h2[idx, dim]
        c2[idx, dim] += h2[idx, dim]
        centres = np.concatenate([c1, c2])
        halves = np.concatenate([h2, h2])
    return (kind, a, b, True, fmin, used)


def main():
    ss, sb = pm.derive_pairs_unpruned()
    tasks = [("ss", int(a), int(b)) for a, b in ss] + [("sb", int(a), int(k)) for a, k in sb]
    t = time.time()
    with Pool(min(8, os.cpu_count() or 1)) as pool:
        res = pool.map(certify, tasks, chunksize=4)
    never_ss = sorted([a, b] for kind, a, b, ok, _, _ in res if ok and kind == "ss")
    never_sb = sorted([a, b] for kind, a, b, ok, _, _ in res if ok and kind == "sb")
    budget = sum(1 for r in res if not r[3] and r[5] > MAX_BOXES)
    out = {
        "margin": MARGIN, "max_boxes": MAX_BOXES,
        "model_fingerprint": pm.model_fingerprint(),
        "never_ss": never_ss, "never_sb": never_sb,
        "stats": {"pairs_in": len(tasks), "certified": len(never_ss) + len(never_sb), "gave_up_on_budget": budget,
                  "seconds": round(time.time() - t, 1)},
    }
    path = os.path.join(ROOT, "rbe550_final_project_b200", "data", "never_collide.json")
    os.makedirs(os.path.dirname(path), exist_ok=True)
    with open(path, "w") as fh:
        json.dump(out, fh)
    print(out["stats"], "ss", len(never_ss), "of", len(ss), "sb", len(never_sb), "of", len(sb))


if __name__ == "__main__":
    main()
