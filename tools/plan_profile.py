"""Where does plan_path's wall time go? (developer tool)"""
import sys, os, time, json, io, contextlib, logging
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from rbe550_final_project_b200 import panda_model as pm, scenes as sc
from rbe550_final_project_b200.planning import PlannerInterface
from rbe550_final_project_b200.sim_stub import create_scene
from rbe550_final_project_b200.validity import PandaValidity
from rbe550_final_project_b200.pathutil import interpolate
logging.getLogger("panda_validity.planning").setLevel(logging.ERROR)
goals = json.load(open(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "goal_configs.json")))
goal = np.array(goals["goal1_scattered"]["approach_r"]["q"])
scene, franka, _ = create_scene("goal1_scattered")
franka.set_qpos(pm.Q_SAFE_HOME)
pv = PandaValidity(0)
planner = PlannerInterface(franka, scene, validity=pv)
def T(fn, n=200):
    for _ in range(5): fn()
    t = time.perf_counter()
    for _ in range(n): fn()
    return (time.perf_counter() - t) / n * 1e6
with contextlib.redirect_stdout(io.StringIO()):
    full = T(lambda: planner.plan_path(qpos_goal=goal, num_waypoints=150, timeout=10.0))
print(f"plan_path total          {full:8.1f} us")
print(f"refresh_scene            {T(planner.refresh_scene):8.1f} us")
sg = np.stack([pm.Q_SAFE_HOME, goal]).astype(np.float32)
print(f"check_states_host(2)     {T(lambda: pv.check_states_host(sg)):8.1f} us")
for rep in (1, 32):
    print(f"rrtc_batch replicas={rep:2d}   {T(lambda: pv.rrtc_batch(sg[0:1], sg[1:2], replicas=rep, max_path=256)):8.1f} us")
path = np.stack([pm.Q_SAFE_HOME, goal])
print(f"interpolate(150)         {T(lambda: interpolate(path, 150)):8.1f} us")
p150 = interpolate(path, 150)
print(f"150 torch tensors        {T(lambda: [torch.tensor(p, dtype=torch.float32) for p in p150]):8.1f} us")
print(f"one tensor + unbind      {T(lambda: list(torch.from_numpy(p150.astype(np.float32)).unbind(0))):8.1f} us")
