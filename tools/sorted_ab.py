"""Same-box A/B of the state kernel's culling modes (1 = per-lane culling, 2 = tile-sorted + culling): timing and
bit-identity of the verdict words, device and host entry points.  Developer tool."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from rbe550_final_project_b200 import panda_model as pm, scenes as sc
from rbe550_final_project_b200.validity import PandaValidity, soa_from_aos

n = (1 << 21) + 77
rng = np.random.default_rng(0)
q = rng.uniform(pm.Q_LOWER, pm.Q_UPPER, size=(n, 9)).astype(np.float32)
q[:, 7:] = 0.04
pv = PandaValidity(0)
A, B, q9 = soa_from_aos(torch.as_tensor(q, device="cuda"))
for scene in ("goal1_scattered", "goal3_tower", "goal4_task1_pentagon"):
    pv.set_scene(sc.FIXTURES[scene]())
    ref = None
    for mode in (1, 2, 1, 2):
        pv.set_culling(mode)
        out = torch.zeros((n + 31) // 32, dtype=torch.int32, device="cuda")
        for _ in range(3):
            pv.check_states((A, B, q9), out=out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            pv.check_states((A, B, q9), out=out)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        w = out.cpu().numpy()
        same = "ref" if ref is None else str(bool((w == ref).all()))
        ref = w if ref is None else ref
        host = pv.check_states_host(q[: 300_001])
        same_host = bool((host.view(np.uint32)[:-1] == ref.view(np.uint32)[: 300_001 // 32]).all())
        print(f"{scene:22s} mode={mode} {n / ms / 1e6:7.3f} G checks/s  identical={same} host_identical={same_host}")
pv.set_culling(1)
# config-5 sweep (device-generated configurations)
pv.set_scene(sc.goal1_scattered())
ns = 20_000_003
ref = None
for mode in (1, 2, 1, 2):
    pv.set_culling(mode)
    for _ in range(2):
        bits, cnt = pv.sweep(0, ns, 20251212)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        bits, cnt = pv.sweep(0, ns, 20251212)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    w = bits.cpu().numpy()
    same = "ref" if ref is None else str(bool((w == ref).all()))
    ref = w if ref is None else ref
    print(f"sweep mode={mode} {ns / ms / 1e6:7.3f} G checks/s n_valid={int(cnt.item())} identical={same}")
pv.set_culling(2)
