"""Offline estimate (numpy + the C oracle's margins): how many listed motions of config 3 would pass their COARSE round
(every second state) with every self-collision / plane test inflated by one step of travel, so that the fine round could
be skipped?  Developer probe for the second-tier certificate."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np
from rbe550_final_project_b200 import panda_model as pm, scenes as sc
from oracle.c_oracle import COracle
from edge_cert_model import clearances, reach_bounds  # noqa (runs its main once on import; small)

scene_name = "goal4_task1_pentagon"
n_e = int(os.environ.get("N", "6000")); nd = 64
rng = np.random.default_rng(3)
qa = rng.uniform(pm.Q_LOWER, pm.Q_UPPER, size=(n_e, 9)); qa[:, 7:] = 0.04
qb = np.clip(qa + 0.3 * rng.standard_normal((n_e, 9)), pm.Q_LOWER, pm.Q_UPPER); qb[:, 7:] = 0.04
Rj = np.asarray(pm.motion_reach_bounds()) if hasattr(pm, "motion_reach_bounds") else reach_bounds()
Rj = np.asarray(Rj, float).ravel()[:7]
dstep = (np.abs(qb - qa)[:, :7] * Rj[None]).sum(1) / nd + np.abs(qb - qa)[:, 7:].sum(1) / nd
ks = nd - 1 - 4 * np.arange(16)
q16 = qa[:, None, :] + (ks / nd)[None, :, None] * (qb - qa)[:, None, :]
pl, se, scn = (x.reshape(n_e, 16) for x in clearances(q16.reshape(-1, 9), scene_name))
d = (dstep * 2)[:, None]
seek = d[:, 0] <= 0.10
u_self = (pl <= d) | (se <= d) | ~seek[:, None]; u_scene = (scn <= d) | ~seek[:, None]
cert = ~(u_self | u_scene).any(1)
cls1 = ~cert & ~u_scene.any(1)
print(f"certified {cert.mean():.3f} class 1 {cls1.mean():.3f} class 2 {(~cert & ~cls1).mean():.3f}; step travel mean {dstep.mean():.4f} p90 {np.quantile(dstep, .9):.4f}")
orc = COracle(pm.model_arrays(), "f32")
empty = {"obb": np.zeros((0, 16)), "table_z": float(sc.FIXTURES[scene_name]().table_z)}
full = sc.FIXTURES[scene_name]().as_oracle_scene()
t = np.arange(1, nd + 1) / nd
qs = qa[:, None, :] + t[None, :, None] * (qb - qa)[:, None, :]
m_self = orc.state_margin(qs.reshape(-1, 9), empty).reshape(n_e, nd)      # self-collision + plane margins, no scene
m_full = orc.state_margin(qs.reshape(-1, 9), full).reshape(n_e, nd)
valid = (m_full > 0).all(1)
even = m_self[:, 1::2]   # states 2, 4, ..., 64  (round 0 of the validator)
for mult, name in ((1.0, "1 step"),):
    ok0 = (even > (dstep * mult)[:, None]).all(1)
    print(f"class-1 motions: valid {valid[cls1].mean():.3f}; coarse round passes with slack {name}: {ok0[cls1].mean():.3f} "
          f"(of the valid ones {ok0[cls1 & valid].mean():.3f})")
    c2 = ~cert & ~cls1
    print(f"class-2 motions: valid {valid[c2].mean():.3f}; self part passes with slack: {ok0[c2].mean():.3f}")
# a quarter-stride variant: states 4, 8, ..., 64 with slack 2 steps, then nothing else
q4 = m_self[:, 3::4]
ok4 = (q4 > (2 * dstep)[:, None]).all(1)
print(f"class-1: stride-4 round with 2 steps of slack passes {ok4[cls1].mean():.3f}")
for mult in (1.0, 1.5, 2.0):
    ok = (even > (dstep * mult)[:, None]).all(1)
    print(f"class-1: coarse round passes with {mult} steps of slack: {ok[cls1].mean():.3f}")
for fixed in (0.02, 0.025, 0.03):
    ok = (even > fixed).all(1) & (dstep <= fixed)
    print(f"class-1: fixed slack {fixed}: {ok[cls1].mean():.3f} (eligible {np.mean(dstep[cls1] <= fixed):.3f})")
# class 2: how many of the 16 coarse cells are unclear for the SCENE-level test (and the plane) alone?
c2 = ~cert & ~cls1 & seek
us = (scn <= d) | (pl <= d)
k = us[c2].sum(1)
print(f"class 2 (certificate sought): scene/plane-unclear cells per motion: mean {k.mean():.2f} of 16; "
      f"P(<= 4) {np.mean(k <= 4):.3f}  P(<= 8) {np.mean(k <= 8):.3f}; scene alone: mean {(scn <= d)[c2].sum(1).mean():.2f}")
