"""Small driver for compute-sanitizer (memcheck / racecheck / synccheck): one launch each of the sorted state kernel, the
motion-validator pipeline (certificate pass + list kernels), the sorted sweep kernel and the batched planner, at sizes a
sanitizer run finishes in a minute or two (ragged: the tail paths run too).  On the GPU box:
    compute-sanitizer --tool racecheck --racecheck-report all python tools/sanitize.py
    compute-sanitizer --tool memcheck python tools/sanitize.py
(compute-sanitizer is closed on the gpurun pool of this project -- rc 86 -- so the committed evidence for the barrier-free state
loop is the argument in DESIGN.md 4.1b: every staging slot is read and re-filled by the thread that owns it.)  Without a
sanitizer the script is a 5-second smoke run of the four pipelines at ragged sizes.  Developer tool."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from rbe550_final_project_b200 import panda_model as pm, scenes as sc  # noqa: E402
from rbe550_final_project_b200.validity import PandaValidity, soa_from_aos  # noqa: E402

what = sys.argv[1:] or ["state", "edges", "sweep", "rrtc"]
pv = PandaValidity(0)
rng = np.random.default_rng(7)


def configs(n):
    q = rng.uniform(pm.Q_LOWER, pm.Q_UPPER, size=(n, 9)).astype(np.float32)
    q[:, 7:] = 0.04
    return q


if "state" in what:
    pv.set_scene(sc.FIXTURES["goal1_scattered"]())
    n = 148 * 512 * 2 + 777  # two chunks per block and a ragged tail
    A, B, _ = soa_from_aos(torch.as_tensor(configs(n), device="cuda"))
    bits = pv.check_states((A, B))
    print("state", n, int(bits.to(torch.int64).sum().item()))
if "edges" in what:
    pv.set_scene(sc.FIXTURES["goal4_task1_pentagon"]())
    n = 40000 + 13
    qa = configs(n)
    qb = np.clip(qa + rng.normal(0, 0.3, qa.shape), pm.Q_LOWER, pm.Q_UPPER).astype(np.float32)
    qb[:, 7:] = 0.04
    out = pv.check_edges(soa_from_aos(torch.as_tensor(qa, device="cuda")), soa_from_aos(torch.as_tensor(qb, device="cuda")), n_steps=64)
    print("edges", n, int(out.to(torch.int64).sum().item()))
if "sweep" in what:
    pv.set_scene(sc.FIXTURES["goal1_scattered"]())
    r = pv.sweep(0, 148 * 512 * 2 + 555, 20251212)
    print("sweep", int((r[0] if isinstance(r, tuple) else r).to(torch.int64).sum().item()))
if "rrtc" in what:
    pv.set_scene(sc.FIXTURES["goal3_tower"]())
    q = configs(4000)
    ok = np.unpackbits(pv.check_states_host(q).view(np.uint8), bitorder="little")[:4000].astype(bool)
    v = q[ok][:128]
    res = pv.rrtc_batch(v[:64], v[64:128], max_iters=200)
    print("rrtc", [getattr(x, "shape", x) for x in (res if isinstance(res, tuple) else (res,))][:3])
torch.cuda.synchronize()
print("ok")
