"""Offline estimate (numpy) of how many config-3 edges a Lipschitz certificate on the culls would finish after the
coarse round: every coarse state's cull tests (plane, self-collision link-pair balls, scene-level test) clear by more
than the largest displacement any robot point can make within `steps` interpolation steps.  Developer probe."""
import os, re, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from rbe550_final_project_b200 import panda_model as pm, scenes as sc
from cull_model import fk

def clearances(q, scene_name):
    """per state: min over cull tests of (distance - threshold), split plane / self / scene"""
    n = q.shape[0]
    R, p = fk(q)
    sl = pm.SPHERE_LINK
    cen = p[:, sl] + np.einsum("nsij,sj->nsi", R[:, sl], pm.SPHERE_CENTER)
    bc = p[:, pm.BOX_LINK] + np.einsum("nkij,kj->nki", R[:, pm.BOX_LINK], pm.BOX_CENTER)
    snap = sc.FIXTURES[scene_name]()
    obb = np.asarray(snap.obb, float)
    oc, oh = obb[:, :3], obb[:, 3:6]
    ext = np.einsum("bij,bj->bi", np.abs(obb[:, 6:15].reshape(-1, 3, 3)), oh)
    lo, hi = (oc - ext).min(0), (oc + ext).max(0)
    tz = snap.table_z
    rad = pm.SPHERE_RADIUS
    m = sl != 0
    plane = (cen[:, m, 2] - rad[None, m]).min(1) - tz
    Rb = R[:, pm.BOX_LINK]
    extb = np.einsum("nkj,kj->nk", np.abs(Rb[:, :, 2, :]), pm.BOX_HALF)
    plane = np.minimum(plane, (bc[:, :, 2] - extb).min(1) - tz)
    bbr = pm.BOX_BOUND_RADIUS
    grip_r = bbr[0] + 2 * pm.CULL_SLACK
    hdr = pm.header_text()
    def rows(name):
        mm = re.search(r"#define %s\(\w+(?:, \w+)*\) \\\n((?:.*\\\n)+)" % name, hdr)
        return [tuple(float(x.rstrip("f")) for x in re.findall(r"[-+0-9.e]+f?", l[l.index("(") + 1:l.rindex(")")]))
                for l in mm.group(1).strip().split("\n") if "(" in l]
    selfc = np.full(n, 1e9)
    for la, lb, ca, cb, c2 in rows("PV_SS_LINKPAIRS"):
        selfc = np.minimum(selfc, np.linalg.norm(cen[:, int(ca)] - cen[:, int(cb)], axis=1) - np.sqrt(c2))
    for la, ca, c0, c1, c2, rla in rows("PV_SBH_LINKS"):
        ca = int(ca)
        if c0 > 0:
            selfc = np.minimum(selfc, np.linalg.norm(cen[:, ca] - bc[:, 0], axis=1) - (rla + grip_r))
        else:
            selfc = np.minimum(selfc, np.linalg.norm(cen[:, ca] - bc[:, 1], axis=1) - np.sqrt(c1))
            selfc = np.minimum(selfc, np.linalg.norm(cen[:, ca] - bc[:, 2], axis=1) - np.sqrt(c2))
    scene = np.full(n, 1e9)
    for (l, cs, br) in pm.link_groups():
        if l == 0: continue
        dd = np.maximum(np.maximum(lo - cen[:, cs], cen[:, cs] - hi), 0).max(1)   # L-inf distance to the bounds
        scene = np.minimum(scene, dd - (br + pm.CULL_SLACK))
    ddg = np.maximum(np.maximum(lo - bc[:, 0], bc[:, 0] - hi), 0).max(1)
    scene = np.minimum(scene, ddg - grip_r)
    return plane, selfc, scene

def reach_bounds():
    """R_j: triangle-inequality bound on the distance from joint j's origin to any robot point distal to it"""
    off = [np.linalg.norm(pm.BODY_POS[i]) for i in range(11)]   # bodies 0..10: link0..7, hand, fingers
    # local extents per body: sphere centres (+0: ball tests use centres), box corners
    loc = np.zeros(11)
    for i in range(len(pm.SPHERE_LINK)):
        loc[pm.SPHERE_LINK[i]] = max(loc[pm.SPHERE_LINK[i]], np.linalg.norm(pm.SPHERE_CENTER[i]))
    for k in range(3):
        loc[pm.BOX_LINK[k]] = max(loc[pm.BOX_LINK[k]], np.linalg.norm(np.abs(pm.BOX_CENTER[k]) + pm.BOX_HALF[k]))
    Rj = []
    for j in range(1, 8):   # joint j moves body j (link j)
        best = 0.0; acc = 0.0
        for b in range(j, 11):
            if b > j: acc += off[b] + (0.04 if b >= 9 else 0)
            if b == 10: acc -= off[9] + 0.04   # fingers are siblings
            best = max(best, acc + loc[b])
        Rj.append(best)
    return np.array(Rj)

def main():
    scene = sys.argv[1] if len(sys.argv) > 1 else "goal4_task1_pentagon"
    n_e = int(sys.argv[2]) if len(sys.argv) > 2 else 4000
    rng = np.random.default_rng(1)
    qa = rng.uniform(pm.Q_LOWER, pm.Q_UPPER, size=(n_e, 9)); qa[:, 7:] = 0.04
    qb = np.clip(qa + 0.3 * rng.standard_normal((n_e, 9)), pm.Q_LOWER, pm.Q_UPPER)
    Rj = reach_bounds()
    print("R_j", np.round(Rj, 3))
    nd = 64
    dstep = (np.abs(qb - qa)[:, :7] * Rj[None]).sum(1) / nd + np.abs(qb - qa)[:, 7:].sum(1) / nd
    print("delta per step: mean %.4f  p90 %.4f" % (dstep.mean(), np.quantile(dstep, 0.9)))
    for stride in (2, 4):
        ks = np.arange(nd, 0, -stride)
        t = ks / nd
        q = qa[:, None, :] + t[None, :, None] * (qb - qa)[:, None, :]
        pl, se, scn = clearances(q.reshape(-1, 9), scene)
        pl, se, scn = (x.reshape(n_e, -1).min(1) for x in (pl, se, scn))
        for name, steps in (("one-sided", stride - 1), ("two-sided", stride // 2)):
            d = dstep * steps
            cert = (pl > d) & (se > d) & (scn > d)
            print(f"stride {stride} {name}: certified {cert.mean():.3f}  (plane {np.mean(pl > d):.3f} self {np.mean(se > d):.3f} scene {np.mean(scn > d):.3f}); "
                  f"cull-free at zero slack {np.mean((pl > 0) & (se > 0) & (scn > 0)):.3f}")



def cells():
    """per uncertified motion: how many of the 16 coarse cells are unclear (self/plane vs scene)"""
    scene = sys.argv[1] if len(sys.argv) > 1 else "goal4_task1_pentagon"
    n_e = int(sys.argv[2]) if len(sys.argv) > 2 else 4000
    rng = np.random.default_rng(1)
    qa = rng.uniform(pm.Q_LOWER, pm.Q_UPPER, size=(n_e, 9)); qa[:, 7:] = 0.04
    qb = np.clip(qa + 0.3 * rng.standard_normal((n_e, 9)), pm.Q_LOWER, pm.Q_UPPER)
    Rj = reach_bounds(); nd = 64; s = 4; h = 2
    dstep = (np.abs(qb - qa)[:, :7] * Rj[None]).sum(1) / nd + np.abs(qb - qa)[:, 7:].sum(1) / nd
    ks = nd - 1 - s * np.arange(16)
    q = qa[:, None, :] + (ks / nd)[None, :, None] * (qb - qa)[:, None, :]
    pl, se, scn = (x.reshape(n_e, 16) for x in clearances(q.reshape(-1, 9), scene))
    d = (dstep * h)[:, None]
    ok_d = (d[:, 0] <= float(os.environ.get("MAXDL", "0.05")))
    u_self = (pl <= d) | (se <= d); u_scene = scn <= d
    u_self[~ok_d] = True; u_scene[~ok_d] = True
    cert = ~(u_self | u_scene).any(1)
    cls1 = ~cert & ~u_scene.any(1); cls2 = ~cert & u_scene.any(1)
    print(f"certified {cert.mean():.3f}  class 1 {cls1.mean():.3f}  class 2 {cls2.mean():.3f}  (no certificate sought {np.mean(~ok_d):.3f})")
    print(f"class 1: unclear cells per motion {u_self[cls1].sum(1).mean():.2f} of 16 -> states {4 * u_self[cls1].sum(1).mean():.1f} of 64; "
          f"P(<= 8 cells, one round) {np.mean(u_self[cls1].sum(1) <= 8):.3f}")
    u = (u_self | u_scene)[cls2]
    print(f"class 2: unclear cells per motion {u.sum(1).mean():.2f} of 16; P(<= 8 cells) {np.mean(u.sum(1) <= 8):.3f}")


if __name__ == "__main__":
    main()
    cells()
