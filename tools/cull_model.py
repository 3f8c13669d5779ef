"""Offline model (numpy, no GPU) of the WARP-LEVEL cost of the per-lane culls in pv_check_config for different visiting
orders of a batch.  A block of tests runs for a whole warp as soon as ONE of its 32 lanes passes the cull in front of it,
so the cost depends on how alike the 32 configurations of a warp are.  Developer tool: it chose the scene-level cull and
the sort key of the sorted kernels, and it predicted (correctly) that larger sort domains and table-driven keys buy
nothing (profiles/r1_notes.md).

What it does: FK of n random configurations (its own small numpy FK from panda_model's body chain), the same culls as the
kernel (link-group balls vs scene boxes, the scene-level test against the padded bounds of all boxes, the self-collision
link-pair culls parsed from csrc/panda_model_gen.h), then for a visiting order: which blocks each warp of 32 consecutive
configurations has to run, priced with rough per-block instruction counts.  Reported per order: mean warp cost of the
scene section, of the self-collision section, and the mean over lockstep iterations of the most expensive of their 16 warps
(measured afterwards: time follows the MEAN, not that maximum).

usage: python tools/cull_model.py [scene] [n]"""
import os, re, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from rbe550_final_project_b200 import panda_model as pm, scenes as sc

# rough warp-instruction prices (from tools/attribute_sass.py on the r1m / r1o kernels)
C_LOAD, C_CULL, C_SPH, C_GRIP = 39, 9, 15, 60   # per scene box: loads, one group cull, one sphere test, gripper path
C_SCENE_TEST = 40                               # the scene-level test in front of the box loop
C_SS_PAIR, C_SBH_XFORM, C_SBH_BOX = 6, 12, 13   # sphere-sphere pair; hand-frame transform; sphere vs one gripper box
C_LP_CULL, C_LB_CULL = 8, 10                    # the culls in front of the self-collision blocks (always run)


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


class Model:
    def __init__(self, scene_name, n, seed=0):
        self.n = n
        rng = np.random.default_rng(seed)
        q = rng.uniform(pm.Q_LOWER, pm.Q_UPPER, size=(n, 9)); q[:, 7:] = 0.04
        self.q = q
        R, p = fk(q)
        sl = pm.SPHERE_LINK
        cen = p[:, sl] + np.einsum("nsij,sj->nsi", R[:, sl], pm.SPHERE_CENTER)
        bc = p[:, pm.BOX_LINK] + np.einsum("nkij,kj->nki", R[:, pm.BOX_LINK], pm.BOX_CENTER)
        obb = np.asarray(sc.FIXTURES[scene_name]().obb, float)
        oc, oh, obr = obb[:, :3], obb[:, 3:6], obb[:, 15]
        ext = np.einsum("bij,bj->bi", np.abs(obb[:, 6:15].reshape(-1, 3, 3)), oh)
        lo, hi = (oc - ext).min(0), (oc + ext).max(0)
        groups = pm.link_groups()
        link_reach, box_reach = pm.static_reach()
        dist0 = np.linalg.norm(oc - (np.asarray(pm.BASE_LIFT) + np.array([0, 0, 0.333])), axis=1) - obr
        bbr = pm.BOX_BOUND_RADIUS
        d1 = np.linalg.norm(bc[:, 1] - bc[:, 0], axis=1); d2 = np.linalg.norm(bc[:, 2] - bc[:, 0], axis=1)
        grip_r = np.maximum(bbr[0], np.maximum(d1 + bbr[1], d2 + bbr[2])) + 2 * pm.CULL_SLACK
        # ---- scene section -------------------------------------------------------------------------------------
        B, G = len(oc), len(groups)
        self.nsph = np.array([int((sl == l).sum()) for l, _, _ in groups])
        self.near = np.zeros((n, B, G), bool); self.stat = np.zeros((B, G), bool)
        aabb_g = np.zeros((n, G), bool)
        for gi, (l, cs, br) in enumerate(groups):
            self.near[:, :, gi] = np.linalg.norm(cen[:, cs, None, :] - oc[None], axis=2) < br + pm.CULL_SLACK + obr[None]
            self.stat[:, gi] = (dist0 < link_reach[l] + 1e-3) if l > 0 else False
            dd = np.linalg.norm(np.maximum(np.maximum(lo - cen[:, cs], cen[:, cs] - hi), 0), axis=1)
            aabb_g[:, gi] = (dd < br + pm.CULL_SLACK) & self.stat[:, gi].any()
        self.gnear = np.linalg.norm(bc[:, 0, None, :] - oc[None], axis=2) < grip_r[:, None] + obr[None]
        self.gstat = dist0 < box_reach.max() + 1e-3
        ddg = np.linalg.norm(np.maximum(np.maximum(lo - bc[:, 0], bc[:, 0] - hi), 0), axis=1)
        self.scene_test = aabb_g.any(1) | (ddg < grip_r)  # the kernel's scene-level test (ball form)
        self.need_scene = (self.near & self.stat[None]).any(axis=(1, 2)) | (self.gnear & self.gstat[None]).any(1)
        wrist = p[:, 5]
        self.wrist_dist = np.linalg.norm(np.maximum(np.maximum(lo - wrist, wrist - hi), 0), axis=1)
        # ---- self-collision section: culls as generated into the header --------------------------------------------
        hdr = pm.header_text()

        def rows(name):
            m = re.search(r"#define %s\(\w+(?:, \w+)*\) \\\n((?:.*\\\n)+)" % name, hdr)
            return [tuple(float(x.rstrip("f")) for x in re.findall(r"[-+0-9.e]+f?", l[l.index("(") + 1:l.rindex(")")]))
                    for l in m.group(1).strip().split("\n") if "(" in l]

        def count(name, what):
            return re.search(r"#define %s\(.*\) \\\n((?:.*\\\n)+)" % name, hdr).group(1).count(what)

        passes, costs, names = [], [], []
        for la, lb, ca, cb, c2 in rows("PV_SS_LINKPAIRS"):
            passes.append(((cen[:, int(ca)] - cen[:, int(cb)]) ** 2).sum(1) < c2)
            costs.append(count("PV_SS_PAIRS_%d_%d" % (la, lb), "X(") * C_SS_PAIR); names.append("SS%d-%d" % (la, lb))
        for la, ca, c0, c1, c2, rla in rows("PV_SBH_LINKS"):
            ca = int(ca)
            if c0 > 0:
                pas = np.linalg.norm(cen[:, ca] - bc[:, 0], axis=1) < rla + grip_r
            else:
                pas = (((cen[:, ca] - bc[:, 1]) ** 2).sum(1) < c1) | (((cen[:, ca] - bc[:, 2]) ** 2).sum(1) < c2)
            passes.append(pas)
            costs.append(count("PV_SBH_%d" % la, "S(") * C_SBH_XFORM + count("PV_SBH_%d" % la, "B(") * C_SBH_BOX)
            names.append("SBH%d" % la)
        self.self_pass = np.stack(passes, 1); self.self_cost = np.array(costs); self.self_names = names
        self.self_fixed = len(rows("PV_SS_LINKPAIRS")) * C_LP_CULL + len(rows("PV_SBH_LINKS")) * C_LB_CULL

    def feat(self, j):
        return (self.q[:, j] - pm.Q_LOWER[j]) / (pm.Q_UPPER[j] - pm.Q_LOWER[j])

    def key(self, edges=(0.20, 0.35, 0.50), q3_bins=16, q5_bins=4):
        """the kernel's sort key: wrist-distance class, elbow bin, wrist-flex bin"""
        cls = np.searchsorted(np.array(edges), self.wrist_dist)
        k3 = np.clip((self.feat(3) * q3_bins).astype(int), 0, q3_bins - 1)
        k5 = np.clip((self.feat(5) * q5_bins).astype(int), 0, q5_bins - 1)
        return (cls * q3_bins + k3) * q5_bins + k5

    def tiles(self, key, tile):
        """each tile of `tile` consecutive configurations is sorted by key (stable)"""
        return np.concatenate([np.argsort(key[i:i + tile], kind="stable") + i for i in range(0, self.n, tile)])

    def warp_costs(self, order, scene_level_test=True):
        w = order[: (len(order) // 512) * 512].reshape(-1, 32)
        any_g = (self.near[w] & self.stat[None, None]).any(1); gw = (self.gnear[w] & self.gstat[None, None]).any(1)
        per_box = (C_LOAD + C_CULL * (self.stat.sum(1)[None] + self.gstat[None])
                   + (any_g * self.nsph[None, None] * C_SPH).sum(2) + gw * C_GRIP)
        scene = per_box.sum(1)
        if scene_level_test:
            scene = np.where(self.scene_test[w].any(1), scene, 0.0) + C_SCENE_TEST
        self_ = self.self_fixed + (self.self_pass[w].any(1) * self.self_cost[None]).sum(1)
        return scene, self_

    def report(self, name, order, scene_level_test=True):
        scene, self_ = self.warp_costs(order, scene_level_test)
        tot = scene + self_
        print(f"{name:52s} scene {scene.mean():7.1f}  self {self_.mean():7.1f}  sum {tot.mean():7.1f}  "
              f"max over the 16 warps of an iteration {tot.reshape(-1, 16).max(1).mean():7.1f}")


def main():
    scene_name = sys.argv[1] if len(sys.argv) > 1 else "goal1_scattered"
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 17
    M = Model(scene_name, n)
    print(f"{scene_name}: P(configuration needs any scene-box test) = {M.need_scene.mean():.4f}, "
          f"P(scene-level test passes) = {M.scene_test.mean():.4f}, P(needs any self-collision block) = {M.self_pass.any(1).mean():.3f}")
    print("P(lane passes) per self-collision cull:", {k: round(float(v), 3) for k, v in zip(M.self_names, M.self_pass.mean(0))})
    ident = np.arange(n)
    T = 7085  # one block's share of a 1 Mi batch on 148 SMs
    k3_256 = np.clip((M.feat(3) * 256).astype(int), 0, 255)
    M.report("unsorted, no scene-level test (r1j)", ident, False)
    M.report("tiles of 7085 sorted by q3 x 256, no scene-level test (r1m)", M.tiles(k3_256, T), False)
    M.report("unsorted + scene-level test", ident)
    M.report("q3 x 256 + scene-level test", M.tiles(k3_256, T))
    M.report("class 4 x q3 64", M.tiles(M.key(q3_bins=64, q5_bins=1), T))
    M.report("class 4 x q3 16 x q5 4 (shipped)", M.tiles(M.key(), T))
    M.report("class 2 x q3 32 x q5 4", M.tiles(M.key(edges=(0.35,), q3_bins=32), T))
    for tile in (2048, 4096, 7168, 14336, n):
        M.report(f"shipped key, sort domain {tile}", M.tiles(M.key(), tile))
    mask = (M.self_pass * (1 << np.arange(M.self_pass.shape[1]))[None]).sum(1)
    cls = np.searchsorted(np.array((0.20, 0.35, 0.50)), M.wrist_dist)
    M.report("perfect: class x exact self-collision cull mask", M.tiles(cls * (1 << 15) + mask, n))


if __name__ == "__main__":
    main()
