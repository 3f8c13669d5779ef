"""Minimal driver for ncu: a few launches of the edge kernel (config-3 workload, 1 Mi edges x 64 steps)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from rbe550_final_project_b200 import _cabi, panda_model as pm, scenes as sc
if os.environ.get("PV_LIB"):  # a variant build (tools/build_variant.py)
    _cabi.LIB_PATH = os.path.join(_cabi.CSRC, os.environ["PV_LIB"])
STEPS = int(os.environ.get("PV_STEPS", "64"))
from rbe550_final_project_b200.validity import PandaValidity, soa_from_aos
pv = PandaValidity(0)
pv.set_scene(sc.FIXTURES[os.environ.get("PV_SCENE", "goal4_task1_pentagon")]())
n = 1 << 20
rng = np.random.default_rng(20251212)
qa = rng.uniform(pm.Q_LOWER, pm.Q_UPPER, size=(n, 9)).astype(np.float32); qa[:, 7:] = 0.04
qb = np.clip(qa + rng.normal(0, 0.3, qa.shape), pm.Q_LOWER, pm.Q_UPPER).astype(np.float32); qb[:, 7:] = 0.04
if os.environ.get("PV_PAIRS") == "uniform":
    qb = rng.uniform(pm.Q_LOWER, pm.Q_UPPER, size=(n, 9)).astype(np.float32); qb[:, 7:] = 0.04
if os.environ.get("PV_CULL"):
    pv.set_culling(int(os.environ["PV_CULL"]))
A = soa_from_aos(torch.as_tensor(qa, device="cuda")); B = soa_from_aos(torch.as_tensor(qb, device="cuda"))
out = torch.empty(n // 32, dtype=torch.int32, device="cuda")
for _ in range(3):
    pv.check_edges(A, B, n_steps=STEPS, out=out)
torch.cuda.synchronize()
print("ok")
