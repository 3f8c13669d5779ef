"""Times alternative builds of the library (csrc/libpv_*.so) on the config-5 sweep. Developer tool."""
import sys, os, glob
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from rbe550_final_project_b200 import _cabi, panda_model as pm, scenes as sc
libs = sorted(glob.glob(os.path.join(_cabi.CSRC, "libpv_*.so")))
n = 104_857_600
for lib in libs:
    _cabi._lib = None; _cabi.LIB_PATH = lib
    from rbe550_final_project_b200.validity import PandaValidity
    pv = PandaValidity(0)
    pv.set_scene(sc.goal1_scattered())
    for _ in range(2): pv.sweep(0, n, 7)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): b, c = pv.sweep(0, n, 7)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print(f"{os.path.basename(lib):20s} {ms:.3f} ms {n/ms/1e6:.2f} G checks/s count={int(c.item())}")
    pv.close()
