"""Same-box A/B of alternative builds of the library (csrc/libpv_*.so, made by tools/build_variant.py).  Developer tool.

    python tools/variant_bench.py [state|edge|sweep|host|bench] [more modes ...] [-- bench.py arguments]

  state  1 Mi / 2 Mi resident configurations, three scenes, culling modes 1 and 2   (pv_check_states)
  edge   1 Mi edges, pentagon / goal-1 scene, 64 fixed steps and resolution mode    (pv_check_edges)
  sweep  104 857 600 device-generated configurations                                  (pv_sweep)
  host   pinned AoS rows through the host entry point                                 (pv_check_states_host)
  bench  bench.py itself (rotating batches, the contract workload) per variant

Box-to-box variance is 3..5 %, so only numbers from ONE invocation are comparable.
"""
import glob
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from rbe550_final_project_b200 import _cabi, panda_model as pm, scenes as sc  # noqa: E402

argv = sys.argv[1:]
bench_args = []
if "--" in argv:
    bench_args = argv[argv.index("--") + 1:]
    argv = argv[: argv.index("--")]
modes = argv or ["state"]
libs = sorted(glob.glob(os.path.join(_cabi.CSRC, "libpv_*.so"))) or [_cabi.LIB_PATH]


def handle(lib):
    _cabi._lib = None
    _cabi.LIB_PATH = lib
    from rbe550_final_project_b200.validity import PandaValidity
    return PandaValidity(0)


def ev(fn, iters, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def configs(n, seed):
    rng = np.random.default_rng(seed)
    q = rng.uniform(pm.Q_LOWER, pm.Q_UPPER, size=(n, 9)).astype(np.float32)
    q[:, 7:] = 0.04
    return q, rng


for mode in modes:
    if mode == "bench":
        args = bench_args or ["--steps", "40", "--warmup", "5", "--no-plan", "--no-cpu-baseline", "--no-configs"]
        for lib in libs:
            code = ("import sys, runpy; sys.path.insert(0, %r); from rbe550_final_project_b200 import _cabi; _cabi.LIB_PATH = %r; "
                    "sys.argv = ['bench.py'] + %r; runpy.run_path(%r, run_name='__main__')" % (ROOT, lib, args, os.path.join(ROOT, "bench.py")))
            out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=ROOT)
            line = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
            if not line:
                print(os.path.basename(lib), "FAILED", out.stderr[-300:])
                continue
            d = json.loads(line[-1])
            print(f"bench {os.path.basename(lib):24s} {d['value'] / 1e9:7.3f} G checks/s  {d['ms_per_step']:7.3f} ms/step  "
                  f"e2e {d['e2e']['value'] / 1e9:.3f} G/s", flush=True)
        continue
    from rbe550_final_project_b200.validity import soa_from_aos
    for lib in libs:
        pv = handle(lib)
        name = os.path.basename(lib)
        if mode == "state":
            n = int(os.environ.get("PV_VB_N", 1 << 21))
            q, _ = configs(n, 0)
            A, B, q9 = soa_from_aos(torch.as_tensor(q, device="cuda"))
            out = torch.empty(n // 32, dtype=torch.int32, device="cuda")
            for scene in ("goal1_scattered", "goal3_tower", "goal4_task1_pentagon"):
                pv.set_scene(sc.FIXTURES[scene]())
                for cull in (1, 2):
                    pv.set_culling(cull)
                    ms = ev(lambda: pv.check_states((A, B), out=out), 10, 3)
                    print(f"state {name:22s} {scene:22s} mode={cull} {n / ms / 1e6:7.3f} G checks/s  chk={int(out.sum().item())}", flush=True)
        elif mode == "edge":
            n = 1 << 20
            qa, rng = configs(n, 20251212)
            qb = np.clip(qa + rng.normal(0, 0.3, qa.shape), pm.Q_LOWER, pm.Q_UPPER).astype(np.float32)
            qb[:, 7:] = 0.04
            A = soa_from_aos(torch.as_tensor(qa, device="cuda"))[:2]
            B = soa_from_aos(torch.as_tensor(qb, device="cuda"))[:2]
            out = torch.empty(n // 32, dtype=torch.int32, device="cuda")
            for scene in ("goal4_task1_pentagon", "goal1_scattered"):
                pv.set_scene(sc.FIXTURES[scene]())
                for steps in (64, 0):
                    ms = ev(lambda: pv.check_edges(A, B, n_steps=steps, out=out), 4, 2)
                    print(f"edge  {name:22s} {scene:22s} n_steps={steps:2d} {n / ms / 1e3:8.2f} M edges/s  chk={int(out.sum().item())}", flush=True)
        elif mode == "sweep":
            n = 104_857_600
            pv.set_scene(sc.goal1_scattered())
            res = {}
            ms = ev(lambda: res.update(r=pv.sweep(0, n, 7)), 3, 2)
            print(f"sweep {name:22s} {ms:.3f} ms {n / ms / 1e6:.2f} G checks/s count={int(res['r'][1].item())}", flush=True)
        elif mode == "host":
            n = 1 << 20
            q, _ = configs(n, 0)
            pv.set_scene(sc.goal1_scattered())
            h = torch.from_numpy(q).pin_memory().numpy()
            out = torch.empty(n // 32, dtype=torch.int32).pin_memory().numpy().view(np.uint32)
            for _ in range(5):
                pv.check_states_host(h, out=out)
            t = time.perf_counter()
            for _ in range(50):
                pv.check_states_host(h, out=out)
            dt = (time.perf_counter() - t) / 50
            print(f"host  {name:22s} {dt * 1e3:.3f} ms  {n / dt / 1e9:.3f} G/s  {n * 36 / dt / 1e9:.1f} GB/s", flush=True)
        else:
            raise SystemExit(f"unknown mode {mode}")
        pv.close()
