"""How much of the state kernel is self-collision? (developer probe: same batch with the self-collision flag off)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from rbe550_final_project_b200 import panda_model as pm, scenes as sc
from rbe550_final_project_b200.validity import PandaValidity, soa_from_aos
pv = PandaValidity(0)
n = 1 << 22
rng = np.random.default_rng(0)
q = rng.uniform(pm.Q_LOWER, pm.Q_UPPER, size=(n, 9)).astype(np.float32); q[:, 7:] = 0.04
A, B, _ = soa_from_aos(torch.as_tensor(q, device="cuda"))
out = torch.empty(n // 32, dtype=torch.int32, device="cuda")
def t(fn, it=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it
for scene in ("goal1_scattered", "goal3_tower", "empty"):
    snap = sc.SceneSnapshot(obb=np.zeros((0, 16), np.float32)) if scene == "empty" else sc.FIXTURES[scene]()
    pv.set_scene(snap)
    for self_on in (True, False):
        pv.set_flags(self_on, False)
        ms = t(lambda: pv.check_states((A, B), out=out))
        print(f"{scene:16s} self={self_on!s:5s} {ms:.3f} ms  {n/ms/1e6:.2f} G checks/s")
