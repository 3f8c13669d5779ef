import sys, os, time, glob
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from rbe550_final_project_b200 import _cabi, panda_model as pm, scenes as sc
libs = sorted(glob.glob(os.path.join(_cabi.CSRC, "libpv_*.so"))) or [_cabi.LIB_PATH]
for lib in libs:
    _cabi._lib = None; _cabi.LIB_PATH = lib
    from rbe550_final_project_b200.validity import PandaValidity, unpack_bits
    pv = PandaValidity(0); pv.set_scene(sc.goal3_tower())
    rng = np.random.default_rng(4096)
    cand = rng.uniform(pm.Q_LOWER, pm.Q_UPPER, size=(60000, 9)).astype(np.float32); cand[:, 7:] = 0.04
    ok = unpack_bits(pv.check_states_host(cand), len(cand))
    ok &= pv.fk(torch.as_tensor(cand, device="cuda")).cpu().numpy()[:, 8, 2] > 0.15
    valid = cand[ok]
    kw = dict(max_iters=2000, max_nodes=2048, max_path=128, seed=7, replicas=1, shortcut_passes=2, packed=True)
    for big in (1 << 15, 1 << 18, 1 << 20):
        idx = rng.integers(0, len(valid), (2, big))
        a, b = valid[idx[0]], valid[idx[1]]
        pv.rrtc_batch(a[:1000], b[:1000], **kw)
        ts = []
        for _ in range(3):
            t = time.perf_counter(); r = pv.rrtc_batch(a, b, **kw); ts.append(time.perf_counter() - t)
        print(os.path.basename(lib), big, f"{min(ts)*1e3:.1f} ms  {big/min(ts)/1e6:.2f} M queries/s  success {(r[2]>=2).mean():.3f}")
    pv.close()
