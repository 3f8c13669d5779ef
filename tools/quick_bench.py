"""Quick device-side timing of the validity kernels (developer tool; bench.py is the contract)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from rbe550_final_project_b200 import panda_model as pm, scenes as sc
from rbe550_final_project_b200.validity import PandaValidity, soa_from_aos, unpack_bits

def timeit(fn, iters=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters

pv = PandaValidity(0)
print("fp32 peak TFLOP/s", pv.fp32_peak(4096))
n = 1 << 22
rng = np.random.default_rng(0)
q = rng.uniform(pm.Q_LOWER, pm.Q_UPPER, size=(n, 9)).astype(np.float32); q[:, 7:] = 0.04
A, B, q9 = soa_from_aos(torch.as_tensor(q, device="cuda"))
out = torch.empty(n // 32, dtype=torch.int32, device="cuda")
for name in ("goal1_scattered", "goal4_task1_pentagon", "goal3_tower"):
    pv.set_scene(sc.FIXTURES[name]())
    for cull in (1, 0):
        pv.set_culling(cull)
        ms = timeit(lambda: pv.check_states((A, B, q9), out=out))
        print(f"{name:22s} cull={cull} states: {ms:.3f} ms  {n/ms/1e6:.3f} G checks/s  valid={unpack_bits(out, n).mean():.3f}")
pv.set_culling(1)
pv.set_scene(sc.goal4_task1_pentagon())
ne = 1 << 20
qb = np.clip(q[:ne] + rng.normal(0, 0.3, (ne, 9)), pm.Q_LOWER, pm.Q_UPPER).astype(np.float32); qb[:, 7:] = 0.04
bA, bB, b9 = soa_from_aos(torch.as_tensor(qb, device="cuda"))
aA, aB, a9 = A[:ne].contiguous(), B[:ne].contiguous(), q9[:ne].contiguous()
oe = torch.empty(ne // 32, dtype=torch.int32, device="cuda")
for steps in (64, 0):
    ms = timeit(lambda: pv.check_edges((aA, aB, a9), (bA, bB, b9), n_steps=steps, out=oe), iters=5)
    print(f"edges n_steps={steps}: {ms:.3f} ms  {ne/ms/1e3:.3f} M edges/s valid={unpack_bits(oe, ne).mean():.3f}")
# host entry point
qh = torch.from_numpy(q).pin_memory().numpy()
oh = np.empty(n // 32, dtype=np.uint32)
pv.set_scene(sc.goal1_scattered())
for _ in range(2): pv.check_states_host(qh, out=oh)
t = time.perf_counter(); 
for _ in range(5): pv.check_states_host(qh, out=oh)
dt = (time.perf_counter() - t) / 5
print(f"host e2e (pinned): {dt*1e3:.3f} ms  {n/dt/1e9:.3f} G checks/s")
