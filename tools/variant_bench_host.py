"""Times pv_check_states_host for alternative builds (chunk size / stream count). Developer tool."""
import sys, os, glob, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from rbe550_final_project_b200 import _cabi, panda_model as pm, scenes as sc
libs = sorted(glob.glob(os.path.join(_cabi.CSRC, "libpv_*.so")))
n = 1 << 20
rng = np.random.default_rng(0)
q = rng.uniform(pm.Q_LOWER, pm.Q_UPPER, size=(n, 9)).astype(np.float32); q[:, 7:] = 0.04
h = torch.from_numpy(q).pin_memory().numpy()
out = torch.empty(n // 32, dtype=torch.int32).pin_memory().numpy().view(np.uint32)
for rep in range(2):
    for lib in libs:
        _cabi._lib = None; _cabi.LIB_PATH = lib
        from rbe550_final_project_b200.validity import PandaValidity
        pv = PandaValidity(0); pv.set_scene(sc.goal1_scattered())
        for _ in range(5): pv.check_states_host(h, out=out)
        t = time.perf_counter()
        for _ in range(50): pv.check_states_host(h, out=out)
        dt = (time.perf_counter() - t) / 50
        print(f"{os.path.basename(lib):16s} {dt*1e3:.3f} ms  {n/dt/1e9:.3f} G/s  {n*36/dt/1e9:.1f} GB/s")
        pv.close()
