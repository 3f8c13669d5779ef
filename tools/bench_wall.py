"""Plan-time comparison on a query that needs real tree growth: a wall between start and goal (not a reference scene;
complements BASELINE config 1, whose goal-1 query is solved by the first straight-line attempt)."""
import sys, os, time, json, io, contextlib, logging
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle.c_oracle import COracle
from rbe550_final_project_b200 import panda_model as pm, scenes as sc
from rbe550_final_project_b200.planning import PlannerInterface
from rbe550_final_project_b200.sim_stub import create_scene
from rbe550_final_project_b200.validity import PandaValidity
from rbe550_final_project_b200.pathutil import interpolate
logging.getLogger("panda_validity.planning").setLevel(logging.ERROR)
pv = PandaValidity(0)
wall = sc.make_obb((0.55, 0.0, 0.35), (0.5, 0.04, 0.7))
snap = sc.SceneSnapshot(obb=np.array([wall], dtype=np.float32), names=["wall"], entity_idx=[1])
pv.set_scene(snap)
quat = np.array([[0.0, 1.0, 0.0, 0.0]])
ql, _, _ = pv.ik_batch(np.array([[0.5, 0.3, 0.3]]), quat, pm.Q_SAFE_HOME, n_seeds=128)
qr, _, _ = pv.ik_batch(np.array([[0.5, -0.3, 0.3]]), quat, pm.Q_SAFE_HOME, n_seeds=128)
scene, franka, _ = create_scene("goal1_scattered")
franka.set_qpos(ql[0])
ora = COracle(pm.model_arrays(), "f32")
out = {"query": "hand (0.5, 0.3, 0.3) -> (0.5, -0.3, 0.3) across a 0.5 x 0.04 x 0.7 m wall", "runs": {}}
for replicas in (1, 8, 32, 128):
    planner = PlannerInterface(franka, snap, validity=pv)
    planner.replicas = replicas
    tg, ok, chk = [], 0, []
    for i in range(104):
        planner.rng_seed = 500 + i
        with contextlib.redirect_stdout(io.StringIO()):
            t = time.perf_counter()
            path = planner.plan_path(qpos_goal=qr[0], num_waypoints=150, timeout=10.0)
            dt = time.perf_counter() - t
        if i >= 3:
            tg.append(dt * 1e3); ok += len(path) >= 150; chk.append(planner.last_stats["state_checks"])
    out["runs"][f"gpu_replicas_{replicas}"] = {"p50_ms": float(np.median(tg)), "p95_ms": float(np.percentile(tg, 95)),
                                                 "success": ok / 101, "median_state_checks_of_winner": float(np.median(chk))}
tc, okc, chc = [], 0, []
for i in range(101):
    t = time.perf_counter()
    p, it, ch = ora.rrtc(ql[0], qr[0], snap.as_oracle_scene(), seed=500 + i, search=0, max_path=256)
    if len(p):
        interpolate(p.astype(np.float64), 150)
    tc.append((time.perf_counter() - t) * 1e3); okc += len(p) > 0; chc.append(ch)
out["runs"]["cpu_port_1core"] = {"p50_ms": float(np.median(tc)), "p95_ms": float(np.percentile(tc, 95)), "success": okc / 101,
                                 "median_state_checks": float(np.median(chc))}
print(json.dumps(out))
