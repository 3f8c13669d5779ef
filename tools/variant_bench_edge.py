"""Times alternative builds of the library (csrc/libpv_*.so) on the config-3 edge workload. Developer tool."""
import sys, os, glob
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from rbe550_final_project_b200 import _cabi, panda_model as pm, scenes as sc
libs = sorted(glob.glob(os.path.join(_cabi.CSRC, "libpv_*.so")))
n = 1 << 20
rng = np.random.default_rng(20251212)
qa = rng.uniform(pm.Q_LOWER, pm.Q_UPPER, size=(n, 9)).astype(np.float32); qa[:, 7:] = 0.04
qb = np.clip(qa + rng.normal(0, 0.3, qa.shape), pm.Q_LOWER, pm.Q_UPPER).astype(np.float32); qb[:, 7:] = 0.04
for lib in libs:
    _cabi._lib = None; _cabi.LIB_PATH = lib
    from rbe550_final_project_b200.validity import PandaValidity, soa_from_aos
    pv = PandaValidity(0)
    A = soa_from_aos(torch.as_tensor(qa, device="cuda")); B = soa_from_aos(torch.as_tensor(qb, device="cuda"))
    out = torch.empty(n // 32, dtype=torch.int32, device="cuda")
    for scene in ("goal4_task1_pentagon", "goal1_scattered"):
        pv.set_scene(sc.FIXTURES[scene]())
        for steps in (64, 0):
            for _ in range(2): pv.check_edges(A, B, n_steps=steps, out=out)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(4): pv.check_edges(A, B, n_steps=steps, out=out)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 4
            print(f"{os.path.basename(lib):18s} {scene:22s} n_steps={steps:2d} {n/ms/1e3:8.2f} M edges/s  chk={int(out.sum().item())}")
    pv.close()
