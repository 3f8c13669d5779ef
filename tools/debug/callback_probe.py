import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from rbe550_final_project_b200 import panda_model as pm, scenes as sc
from rbe550_final_project_b200.validity import PandaValidity
pv = PandaValidity(0); pv.set_scene(sc.goal1_scattered())
q = np.asarray(pm.Q_SAFE_HOME, np.float32).reshape(1, 9)
out = np.empty(1, np.uint32)
def T(fn, n=2000):
    for _ in range(50): fn()
    ts = []
    for _ in range(n):
        t = time.perf_counter(); fn(); ts.append(time.perf_counter() - t)
    return np.median(ts) * 1e6, np.percentile(ts, 95) * 1e6
print("is_state_valid            p50 %.1f us p95 %.1f" % T(lambda: pv.is_state_valid(pm.Q_SAFE_HOME)))
print("check_states_host(1)      p50 %.1f us p95 %.1f" % T(lambda: pv.check_states_host(q, out=out)))
q32 = np.repeat(q, 32, 0); out32 = np.empty(1, np.uint32)
print("check_states_host(32)     p50 %.1f us p95 %.1f" % T(lambda: pv.check_states_host(q32, out=out32)))
qa = np.repeat(q, 150, 0); o = np.empty(5, np.uint32)
print("check_edges_host(150)     p50 %.1f us p95 %.1f" % T(lambda: pv.check_edges_host(qa, qa, out=o)))
