"""Same-box A/B of the motion certificates: pv_set_culling(1) (exhaustive) vs pv_set_culling(2) with the first-tier
certificates only (a handle created under PV_EDGE_CERT2=0) vs both tiers (the default).  Developer probe."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from rbe550_final_project_b200 import panda_model as pm, scenes as sc
from rbe550_final_project_b200.validity import PandaValidity

def handle(mode):  # PV_EDGE_CERT2 = 0: a handle without the second-tier certificate pass
    os.environ["PV_EDGE_CERT2"] = str(mode)
    h = PandaValidity(0)
    os.environ.pop("PV_EDGE_CERT2")
    return h
pvs = {m: handle(m) for m in (0, 1)}
pv = pvs[1]
n = 1 << 20
g = torch.Generator(device="cuda"); g.manual_seed(20251212)
lo = torch.tensor(pm.Q_LOWER, dtype=torch.float32, device="cuda"); hi = torch.tensor(pm.Q_UPPER, dtype=torch.float32, device="cuda")
qa = lo + (hi - lo) * torch.rand((n, 9), generator=g, device="cuda"); qa[:, 7:] = 0.04
qb = torch.minimum(torch.maximum(qa + 0.3 * torch.randn((n, 9), generator=g, device="cuda"), lo), hi); qb[:, 7:] = 0.04
qu = lo + (hi - lo) * torch.rand((n, 9), generator=g, device="cuda"); qu[:, 7:] = 0.04
bits = torch.empty(n // 32, dtype=torch.int32, device="cuda")

def ev(fn, iters=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters

for scene in ("goal4_task1_pentagon", "goal1_scattered", "goal3_tower"):
    for h in pvs.values():
        h.set_scene(sc.FIXTURES[scene]())
    for what, b_, ns in (("gauss 64", qb, 64), ("gauss 128", qb, 128), ("gauss res", qb, 0), ("uniform 64", qu, 64), ("uniform res", qu, 0)):
        out = {}
        for mode in (1, 2):
            pv.set_culling(mode)
            ms = ev(lambda: pv.check_edges(qa, b_, n_steps=ns, out=bits))
            out[mode] = (ms, bits.clone())
        same = bool((out[1][1] == out[2][1]).all())
        ms1 = ev(lambda: pvs[0].check_edges(qa, b_, n_steps=ns, out=bits))
        same = same and bool((out[1][1] == bits).all())
        print(f"{scene:22s} {what:12s} exhaustive {n / out[1][0] / 1e3:7.1f}   tier 1 {n / ms1 / 1e3:7.1f}   tiers 1+2 {n / out[2][0] / 1e3:7.1f} M edges/s   identical {same}", flush=True)
pv.set_culling(2)
