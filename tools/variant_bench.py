"""Times alternative builds of the library (csrc/libpv_*.so) on the bench workload. Developer tool."""
import sys, os, glob
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from rbe550_final_project_b200 import _cabi, panda_model as pm, scenes as sc
libs = sorted(glob.glob(os.path.join(_cabi.CSRC, "libpv_*.so")))
n = int(os.environ.get("PV_VB_N", 1 << 21))
rng = np.random.default_rng(0)
q = rng.uniform(pm.Q_LOWER, pm.Q_UPPER, size=(n, 9)).astype(np.float32); q[:, 7:] = 0.04
for lib in libs:
    _cabi._lib = None; _cabi.LIB_PATH = lib
    from rbe550_final_project_b200.validity import PandaValidity, soa_from_aos
    pv = PandaValidity(0)
    A, B, q9 = soa_from_aos(torch.as_tensor(q, device="cuda"))
    out = torch.empty(n // 32, dtype=torch.int32, device="cuda")
    for scene in ("goal1_scattered", "goal3_tower", "goal4_task1_pentagon"):
        pv.set_scene(sc.FIXTURES[scene]())
        for mode in (1, 2):
            pv.set_culling(mode)
            for _ in range(3): pv.check_states((A, B, q9), out=out)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10): pv.check_states((A, B, q9), out=out)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 10
            print(f"{os.path.basename(lib):22s} {scene:18s} mode={mode} {n/ms/1e6:.3f} G checks/s  chk={int(out.sum().item())}")
    pv.close()
