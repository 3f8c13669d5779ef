"""Minimal driver for ncu: a few launches of the state-validity kernel on 1M configs (goal-1 scene)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from rbe550_final_project_b200 import panda_model as pm, scenes as sc
from rbe550_final_project_b200.validity import PandaValidity, soa_from_aos

scene = sys.argv[1] if len(sys.argv) > 1 else "goal1_scattered"
cull = int(sys.argv[2]) if len(sys.argv) > 2 else 2
pv = PandaValidity(0)
pv.set_scene(sc.FIXTURES[scene]())
pv.set_culling(cull)
n = 1 << 20
rng = np.random.default_rng(0)
q = rng.uniform(pm.Q_LOWER, pm.Q_UPPER, size=(n, 9)).astype(np.float32); q[:, 7:] = 0.04
A, B, q9 = soa_from_aos(torch.as_tensor(q, device="cuda"))
out = torch.empty(n // 32, dtype=torch.int32, device="cuda")
for _ in range(4):
    pv.check_states((A, B), out=out)  # symmetric gripper: two float4 planes, as in bench.py
torch.cuda.synchronize()
print("ok", int(out[0].item()))
