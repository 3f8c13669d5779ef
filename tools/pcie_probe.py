import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from rbe550_final_project_b200 import panda_model as pm, scenes as sc
from rbe550_final_project_b200.validity import PandaValidity
n = 1 << 20
h = torch.empty((n, 9), dtype=torch.float32).pin_memory()
d = torch.empty((n, 9), dtype=torch.float32, device="cuda")
for sz in (n, n // 4, n // 16):
    for _ in range(3): d[:sz].copy_(h[:sz], non_blocking=True)
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(20): d[:sz].copy_(h[:sz], non_blocking=True)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t) / 20
    print(f"H2D pinned {sz*36/1e6:.1f} MB: {sz*36/dt/1e9:.1f} GB/s")
pv = PandaValidity(0); pv.set_scene(sc.goal1_scattered())
rng = np.random.default_rng(0)
q = rng.uniform(pm.Q_LOWER, pm.Q_UPPER, size=(n, 9)).astype(np.float32); q[:, 7:] = 0.04
h.copy_(torch.from_numpy(q))
out = torch.empty(n // 32, dtype=torch.int32).pin_memory().numpy().view(np.uint32)
hq = h.numpy()
for _ in range(3): pv.check_states_host(hq, out=out)
t = time.perf_counter()
for _ in range(20): pv.check_states_host(hq, out=out)
dt = (time.perf_counter() - t) / 20
print(f"check_states_host 1M: {dt*1e3:.3f} ms  {n/dt/1e9:.3f} G/s  ({n*36/dt/1e9:.1f} GB/s H2D equivalent)")
