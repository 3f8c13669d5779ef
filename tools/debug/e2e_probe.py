import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from rbe550_final_project_b200 import panda_model as pm, scenes as sc
from rbe550_final_project_b200.validity import PandaValidity
pv = PandaValidity(0); pv.set_scene(sc.goal1_scattered())
rng = np.random.default_rng(0)
N = 1 << 22
q = rng.uniform(pm.Q_LOWER, pm.Q_UPPER, size=(N, 9)).astype(np.float32); q[:, 7:] = 0.04
h = torch.from_numpy(q).pin_memory(); hq = h.numpy()
out = torch.empty(N // 32, dtype=torch.int32).pin_memory().numpy().view(np.uint32)
d = torch.as_tensor(q, device="cuda")
from rbe550_final_project_b200.validity import soa_from_aos
A, B, q9 = soa_from_aos(d)
for n in (1 << 15, 1 << 17, 1 << 18, 1 << 19, 1 << 20, 1 << 22):
    for _ in range(3): pv.check_states_host(hq[:n], out=out[: n // 32])
    t = time.perf_counter()
    for _ in range(20): pv.check_states_host(hq[:n], out=out[: n // 32])
    dt = (time.perf_counter() - t) / 20
    # device-only kernels on the same data: AoS and SoA
    def ev(fn, it=20):
        fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(it): fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / it
    ms_soa = ev(lambda: pv.check_states((A[:n], B[:n])))
    dd = torch.empty((n, 9), dtype=torch.float32, device="cuda")
    ms_copy = ev(lambda: dd.copy_(h[:n], non_blocking=True))
    print(f"n={n:8d}: host call {dt*1e3:.3f} ms = {n/dt/1e9:.3f} G/s ({n*36/dt/1e9:.1f} GB/s) | plain H2D {ms_copy:.3f} ms ({n*36/ms_copy/1e6:.1f} GB/s) | SoA kernel {ms_soa:.3f} ms")
# rotating pinned buffers (host DRAM, not LLC)
n = 1 << 20
bufs = [torch.from_numpy(q[:n].copy()).pin_memory() for _ in range(8)]
dd = torch.empty((n, 9), dtype=torch.float32, device="cuda")
for k in (1, 2, 4, 8):
    for i in range(8): dd.copy_(bufs[i % k], non_blocking=True)
    torch.cuda.synchronize()
    t = time.perf_counter()
    for i in range(40): dd.copy_(bufs[i % k], non_blocking=True)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t) / 40
    for i in range(4): pv.check_states_host(bufs[i % k].numpy(), out=out[: n // 32])
    t = time.perf_counter()
    for i in range(40): pv.check_states_host(bufs[i % k].numpy(), out=out[: n // 32])
    dc = (time.perf_counter() - t) / 40
    print(f"rotate {k} buffers: plain H2D {n*36/dt/1e9:.1f} GB/s | host call {n/dc/1e9:.3f} G/s ({n*36/dc/1e9:.1f} GB/s)")
# 7-column rows
b7 = [torch.from_numpy(np.ascontiguousarray(q[:n, :7])).pin_memory() for _ in range(8)]
d7 = torch.empty((n, 7), dtype=torch.float32, device="cuda")
for i in range(8): d7.copy_(b7[i], non_blocking=True)
torch.cuda.synchronize()
t = time.perf_counter()
for i in range(40): d7.copy_(b7[i % 8], non_blocking=True)
torch.cuda.synchronize()
dt = (time.perf_counter() - t) / 40
print(f"7-column rows, 8 buffers: plain H2D {n*28/dt/1e9:.1f} GB/s -> ceiling {n/dt/1e9:.3f} G checks/s")
