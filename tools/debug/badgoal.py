import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from rbe550_final_project_b200 import panda_model as pm, scenes as sc
from rbe550_final_project_b200.validity import PandaValidity, unpack_bits
pv = PandaValidity(0)
snap = sc.goal1_scattered(); snap.obb[0, 0:3] = (1.4, 0.0, 0.02)
pv.set_scene(snap)
q = np.array([[0, 1.7, 0, -0.1, 0, 0.5, 0, 0.04, 0.04]], np.float32)
print("host bits", pv.check_states_host(q))
print("margin", pv.state_margins(torch.as_tensor(q, device="cuda"), want_culprit=True))
for mode in (0, 1, 2):
    pv.set_culling(mode)
    print("cull", mode, pv.check_states_host(q), pv.check_states(torch.as_tensor(np.repeat(q, 64, 0), device="cuda")).cpu().numpy())
pv.set_culling(2)
print("rrtc", pv.rrtc_batch(np.asarray([pm.Q_SAFE_HOME], np.float32), q, check_endpoints=True, replicas=1)[1:])
print("rrtc32", pv.rrtc_batch(np.asarray([pm.Q_SAFE_HOME], np.float32), q, check_endpoints=True, replicas=32)[1:])
wp, st = pv.plan_path(pm.Q_SAFE_HOME, q[0], num_waypoints=150)
print("plan", len(wp), st)
