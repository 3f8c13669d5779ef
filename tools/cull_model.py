"""Offline model (numpy, no GPU) of the WARP-LEVEL cost of the per-lane culls in pv_check_config for different
visiting orders of a batch.  A block of tests runs for a whole warp as soon as ONE of its 32 lanes passes the cull in
front of it, so the cost depends on how alike the 32 configurations of a warp are.  Developer tool: it decided the sort
key of pv_state_bits_sorted_kernel (profiles/r1_notes.md).

usage: python tools/cull_model.py [scene] [n]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from rbe550_final_project_b200 import panda_model as pm, scenes as sc


def quat_mat(q):
    w, x, y, z = np.asarray(q, float) / np.linalg.norm(q)
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def fk(q):
    n = q.shape[0]
    R = np.zeros((n, 11, 3, 3)); p = np.zeros((n, 11, 3))
    joint = [None, 0, 1, 2, 3, 4, 5, 6, None, ("p", 7), ("p", 8)]
    for i in range(11):
        Ri = quat_mat(pm.BODY_QUAT[i]); pi = np.asarray(pm.BODY_POS[i], float)
        if pm.PARENT[i] < 0:
            Rw = np.broadcast_to(Ri, (n, 3, 3)).copy(); pw = np.broadcast_to(np.asarray(pm.BASE_LIFT) + pi, (n, 3)).copy()
        else:
            Rp, pp = R[:, pm.PARENT[i]], p[:, pm.PARENT[i]]
            Rw = Rp @ Ri; pw = pp + Rp @ pi
        j = joint[i]
        if isinstance(j, int):
            c, s = np.cos(q[:, j]), np.sin(q[:, j])
            Rz = np.zeros((n, 3, 3)); Rz[:, 0, 0], Rz[:, 0, 1], Rz[:, 1, 0], Rz[:, 1, 1], Rz[:, 2, 2] = c, -s, s, c, 1
            Rw = Rw @ Rz
        elif j is not None:
            pw = pw + Rw[:, :, 1] * q[:, j[1]:j[1] + 1]
        R[:, i], p[:, i] = Rw, pw
    return R, p


def main():
    scene_name = sys.argv[1] if len(sys.argv) > 1 else "goal1_scattered"
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 16
    rng = np.random.default_rng(0)
    q = rng.uniform(pm.Q_LOWER, pm.Q_UPPER, size=(n, 9)); q[:, 7:] = 0.04
    R, p = fk(q)
    sl = pm.SPHERE_LINK
    cen = p[:, sl] + np.einsum("nsij,sj->nsi", R[:, sl], pm.SPHERE_CENTER)  # (n, S, 3)
    bc = p[:, pm.BOX_LINK] + np.einsum("nkij,kj->nki", R[:, pm.BOX_LINK], pm.BOX_CENTER)
    snap = sc.FIXTURES[scene_name]()
    obb = np.asarray(snap.obb, float)  # (B, 16)
    oc, obr = obb[:, :3], obb[:, 15]
    B = oc.shape[0]
    groups = pm.link_groups()
    link_reach, box_reach = pm.static_reach()
    s0 = np.asarray(pm.BASE_LIFT) + np.array([0, 0, 0.333])
    dist0 = np.linalg.norm(oc - s0, axis=1) - obr
    nsph = {l: int((sl == l).sum()) for l in range(8)}
    # per-lane cull outcomes: near[n, b, g]
    near = np.zeros((n, B, len(groups)), bool); stat = np.zeros((B, len(groups)), bool)
    for gi, (l, cs, br) in enumerate(groups):
        d = np.linalg.norm(cen[:, cs, None, :] - oc[None], axis=2)
        near[:, :, gi] = d < br + pm.CULL_SLACK + obr[None]
        stat[:, gi] = (dist0 < link_reach[l] + 1e-3) if l > 0 else True
    d1 = np.linalg.norm(bc[:, 1] - bc[:, 0], axis=1); d2 = np.linalg.norm(bc[:, 2] - bc[:, 0], axis=1)
    bbr = pm.BOX_BOUND_RADIUS
    grip_r = np.maximum(bbr[0], np.maximum(d1 + bbr[1], d2 + bbr[2])) + 2 * pm.CULL_SLACK
    gnear = np.linalg.norm(bc[:, 0, None, :] - oc[None], axis=2) < grip_r[:, None] + obr[None]
    gstat = dist0 < box_reach.max() + 1e-3
    print(f"{scene_name}: B={B}; static reach per group: {stat.sum(0)} of {B}; gripper {gstat.sum()}")
    print("P(lane passes cull) per group (mean over reachable boxes):",
          [round(float(near[:, stat[:, gi], gi].mean()), 3) if stat[:, gi].any() else None for gi in range(len(groups))],
          "gripper", round(float(gnear[:, gstat].mean()), 3))
    anynear = (near & stat[None]).any(axis=(1, 2)) | (gnear & gstat[None]).any(axis=1)
    print(f"P(config needs ANY scene-box test) = {anynear.mean():.3f}")

    C_LOAD, C_CULL, C_SPH, C_GRIP = 39, 9, 15, 60  # warp instructions: box loads, one cull, one sphere test, gripper path

    def cost(order, with_skip):
        """mean warp instructions per 32 configurations spent on the scene-box section"""
        w = order[: (n // 32) * 32].reshape(-1, 32)
        nw = near[w] & stat[None, None]  # (W, 32, B, G)
        any_g = nw.any(axis=1)  # (W, B, G)
        gw = (gnear[w] & gstat[None, None]).any(axis=1)  # (W, B)
        c = np.zeros(w.shape[0])
        per_box = C_LOAD + C_CULL * (stat.sum(1)[None] + gstat[None])  # culls always run
        per_box = per_box + (any_g * np.array([nsph[l] for l, _, _ in groups])[None, None] * C_SPH).sum(2) + gw * C_GRIP
        c = per_box.sum(1)
        if with_skip:  # one scene-level test per warp in front of the whole box loop
            need = any_g.any(axis=(1, 2)) | gw.any(axis=1)
            c = np.where(need, c, 0.0) + 25
        return c.mean()

    ident = np.arange(n)
    k3 = np.clip(((q[:, 3] - pm.Q_LOWER[3]) / (pm.Q_UPPER[3] - pm.Q_LOWER[3]) * 256).astype(int), 0, 255)
    by_q3 = np.concatenate([np.argsort(k3[i:i + 16384], kind="stable") + i for i in range(0, n, 16384)])
    # wrist (link6 origin) distance to the scene's bounding box
    lo = (oc - obr[:, None]).min(0); hi = (oc + obr[:, None]).max(0)
    wr = p[:, 6]
    dd = np.linalg.norm(np.maximum(np.maximum(lo - wr, wr - hi), 0), axis=1)
    for nb_d, nb_3 in ((2, 128), (4, 64), (8, 32), (16, 16)):
        edges = np.quantile(dd, np.linspace(0, 1, nb_d + 1)[1:-1])
        kd = np.searchsorted(edges, dd)
        key = kd * nb_3 + (k3 * nb_3 // 256)
        o = np.concatenate([np.argsort(key[i:i + 16384], kind="stable") + i for i in range(0, n, 16384)])
        print(f"sort (dist class {nb_d} x q3 {nb_3}):  scene cost {cost(o, False):7.1f}   with scene-level skip {cost(o, True):7.1f}")
    # exact class: does the configuration need any scene test (upper bound on what a distance key can deliver)
    key = anynear.astype(int) * 256 + k3
    o = np.concatenate([np.argsort(key[i:i + 16384], kind="stable") + i for i in range(0, n, 16384)])
    print(f"sort (needs-scene x q3 256):  scene cost {cost(o, False):7.1f}   with scene-level skip {cost(o, True):7.1f}")
    print(f"unsorted:            scene cost {cost(ident, False):7.1f}   with skip {cost(ident, True):7.1f}")
    print(f"sorted by q3 (now):  scene cost {cost(by_q3, False):7.1f}   with skip {cost(by_q3, True):7.1f}")


if __name__ == "__main__":
    main()
