import sys, os, time, json, io, contextlib, logging, cProfile, pstats
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from rbe550_final_project_b200 import panda_model as pm
from rbe550_final_project_b200.planning import PlannerInterface
from rbe550_final_project_b200.sim_stub import create_scene
from rbe550_final_project_b200.validity import PandaValidity
goals = json.load(open(os.path.join(os.path.dirname(__file__), "..", "..", "tests", "golden", "goal_configs.json")))
goal = np.array(goals["goal1_scattered"]["approach_r"]["q"])
scene, franka, _ = create_scene("goal1_scattered")
franka.set_qpos(pm.Q_SAFE_HOME)
pv = PandaValidity(0)
planner = PlannerInterface(franka, scene, validity=pv)
sink = io.StringIO()
def run(n):
    ts = []
    for i in range(n):
        sink.seek(0); sink.truncate()
        with contextlib.redirect_stdout(sink):
            t = time.perf_counter()
            planner.plan_path(qpos_goal=goal, num_waypoints=150, timeout=10.0)
            ts.append(time.perf_counter() - t)
    return np.array(ts) * 1e3
run(50)
ts = run(1000)
print(f"p50 {np.median(ts):.4f} ms  p95 {np.percentile(ts,95):.4f}  min {ts.min():.4f}")
st = planner.last_stats
print({k: st[k] for k in ("ms_scene_snapshot", "ms_c_call", "ms_python_rest", "ms_solve", "ms_total", "launches")})
pr = cProfile.Profile()
pr.enable(); run(2000); pr.disable()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(28)
print(s.getvalue()[:6000])
