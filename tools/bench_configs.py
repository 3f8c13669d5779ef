"""Measures the BASELINE.json configs that bench.py does not time (1, 3, 4) and prints one JSON line each.

  config 1  goal1_scattered single plan: p50 plan_path time on the GPU vs the CPU oracle planner (port, 1 core)
  config 3  10 485 760 edges x 64 interpolation steps, finished-pentagon scene (goal4_task1)
  config 4  4096 start/goal pairs, batched RRT-Connect, tall-tower scene (goal3)
Run on a B200:  python tools/bench_configs.py [1|3|4 ...]
"""
import contextlib, io, json, logging, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from rbe550_final_project_b200 import panda_model as pm, scenes as sc
from rbe550_final_project_b200.validity import PandaValidity, soa_from_aos, unpack_bits

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
which = [int(a) for a in sys.argv[1:]] or [1, 3, 4]
pv = PandaValidity(0)


def ev_time(fn, iters, warm=2):
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


if 1 in which:
    from oracle.c_oracle import COracle
    from rbe550_final_project_b200.planning import PlannerInterface
    from rbe550_final_project_b200.sim_stub import create_scene
    from rbe550_final_project_b200.pathutil import interpolate
    logging.getLogger("panda_validity.planning").setLevel(logging.ERROR)
    goals = json.load(open(os.path.join(ROOT, "tests", "golden", "goal_configs.json")))
    ora = COracle(pm.model_arrays(), "f32")
    out = {"config": 1, "what": "single RRT-Connect plan_path, safe_home -> goal, smooth, 150 waypoints; p50 over 101 seeds",
           "cases": {}}
    for scene_name, cases in (("goal1_scattered", ["approach_r", "grasp_r", "approach_c"]),
                              ("goal3_tower", ["approach_top", "approach_r2"])):
        scene, franka, _ = create_scene(scene_name)
        franka.set_qpos(pm.Q_SAFE_HOME)
        planner = PlannerInterface(franka, scene, validity=pv)
        oscene = sc.FIXTURES[scene_name]().as_oracle_scene()
        for case in cases:
            goal = np.array(goals[scene_name][case]["q"])
            tg, tc, okg, okc = [], [], 0, 0
            for i in range(104):
                planner.rng_seed = 100 + i
                with contextlib.redirect_stdout(io.StringIO()):
                    t = time.perf_counter()
                    path = planner.plan_path(qpos_goal=goal, num_waypoints=150, timeout=10.0)
                    dt = time.perf_counter() - t
                t = time.perf_counter()
                p, it, ch = ora.rrtc(pm.Q_SAFE_HOME, goal, oscene, seed=100 + i, search=0, max_path=256)
                if len(p):
                    interpolate(p.astype(np.float64), 150)
                dc = time.perf_counter() - t
                if i >= 3:
                    tg.append(dt * 1e3); tc.append(dc * 1e3)
                    okg += len(path) == 150; okc += len(p) > 0
            out["cases"][f"{scene_name}/{case}"] = {
                "gpu_p50_ms": float(np.median(tg)), "gpu_p95_ms": float(np.percentile(tg, 95)), "gpu_success": okg / 101,
                "cpu_port_p50_ms": float(np.median(tc)), "cpu_port_p95_ms": float(np.percentile(tc, 95)),
                "cpu_port_success": okc / 101}
    out["note"] = ("cpu_port = oracle/rrtc_oracle_impl.h (C, fp32, 1 core, same geometry model) -- NOT Genesis+OMPL, whose "
                   "per-state cost is dominated by Python/Taichi dispatch (SURVEY.md 3.3) and which cannot be installed here")
    print(json.dumps(out))

if 3 in which:
    snap = sc.goal4_task1_pentagon()
    pv.set_scene(snap)
    n = 10_485_760
    rng = np.random.default_rng(20251212)
    res = {"config": 3, "scene": "goal4_task1_pentagon", "n_edges": n, "n_steps": 64, "variants": {}}
    qa = rng.uniform(pm.Q_LOWER, pm.Q_UPPER, size=(n, 9)).astype(np.float32); qa[:, 7:] = 0.04
    for variant in ("gaussian_0.3", "uniform_pairs"):
        if variant == "gaussian_0.3":
            qb = np.clip(qa + rng.normal(0, 0.3, qa.shape), pm.Q_LOWER, pm.Q_UPPER).astype(np.float32)
        else:
            qb = rng.uniform(pm.Q_LOWER, pm.Q_UPPER, size=(n, 9)).astype(np.float32)
        qb[:, 7:] = 0.04
        A = soa_from_aos(torch.as_tensor(qa, device="cuda")); B = soa_from_aos(torch.as_tensor(qb, device="cuda"))
        out_bits = torch.empty(n // 32, dtype=torch.int32, device="cuda")
        ms = ev_time(lambda: pv.check_edges(A, B, n_steps=64, out=out_bits), iters=3, warm=1)
        valid = float(np.unpackbits(out_bits.cpu().numpy().view(np.uint8)).sum()) / n
        res["variants"][variant] = {"ms": ms, "edges_per_s": n / (ms * 1e-3), "state_evals_per_s_upper": 64 * n / (ms * 1e-3),
                                    "valid_fraction": valid}
        del A, B
    # CPU port on a bounded sample (early exit like OMPL's validator), all cores
    from oracle.c_oracle import COracle
    ora = COracle(pm.model_arrays(), "f32")
    k = 20000
    qb = np.clip(qa[:k] + rng.normal(0, 0.3, (k, 9)), pm.Q_LOWER, pm.Q_UPPER).astype(np.float32); qb[:, 7:] = 0.04
    t = time.perf_counter()
    _, cnt = ora.edge_margin(qa[:k], qb, snap.as_oracle_scene(), n_steps=64, early_exit=True, return_count=True)
    dt = time.perf_counter() - t
    res["cpu_port"] = {"edges_per_s": k / dt, "cores": os.cpu_count(), "sample_edges": k, "state_evals": cnt}
    print(json.dumps(res))

if 4 in which:
    snap = sc.goal3_tower()
    pv.set_scene(snap)
    from oracle import panda_oracle as po
    from oracle.c_oracle import COracle
    ora64 = COracle(pm.model_arrays(), "f64")
    rng = np.random.default_rng(4096)
    cand = rng.uniform(pm.Q_LOWER, pm.Q_UPPER, size=(60000, 9)).astype(np.float32); cand[:, 7:] = 0.04
    ok = unpack_bits(pv.check_states_host(cand), len(cand))
    # hand z > 0.15: hand height from the FK kernel
    poses = pv.fk(torch.as_tensor(cand, device="cuda")).cpu().numpy()
    ok &= poses[:, 8, 2] > 0.15
    valid = cand[ok]
    nq = 4096
    starts, goals = valid[:nq], valid[nq:2 * nq]
    res = {"config": 4, "scene": "goal3_tower", "n_queries": nq, "max_iters": 2000, "runs": {}}
    for replicas in (1, 4):
        # warm-up with the full problem: the first call sizes the device arena and the pinned result mirror
        pv.rrtc_batch(starts, goals, max_iters=2000, max_nodes=2048, max_path=128, seed=3, replicas=replicas, shortcut_passes=2)
        t = time.perf_counter()
        paths, plen, iters, checks = pv.rrtc_batch(starts, goals, max_iters=2000, max_nodes=2048, max_path=128, seed=7,
                                                   replicas=replicas, shortcut_passes=2)
        dt = time.perf_counter() - t
        solved = plen > 0
        res["runs"][f"replicas_{replicas}"] = {
            "wall_ms": dt * 1e3, "queries_per_s": nq / dt, "success": float(solved.mean()),
            "iters_p50": float(np.median(iters[solved])), "iters_p95": float(np.percentile(iters[solved], 95)),
            "state_checks_total": int(checks.sum()), "path_len_p50": float(np.median(plen[solved]))}
    # per-query latency distribution: one query per launch (what plan_path does)
    lat = []
    for k in range(200):
        t = time.perf_counter()
        pv.rrtc_batch(starts[k:k + 1], goals[k:k + 1], replicas=32, seed=7)
        lat.append((time.perf_counter() - t) * 1e3)
    res["single_query_launch_ms"] = {"p50": float(np.median(lat)), "p95": float(np.percentile(lat, 95))}
    # CPU port, 1 core, first 200 queries
    ora = COracle(pm.model_arrays(), "f32")
    tc, okc = [], 0
    for k in range(200):
        t = time.perf_counter()
        p, it, ch = ora.rrtc(starts[k], goals[k], snap.as_oracle_scene(), seed=7, search=k)
        tc.append((time.perf_counter() - t) * 1e3); okc += len(p) > 0
    res["cpu_port_1core"] = {"p50_ms": float(np.median(tc)), "p95_ms": float(np.percentile(tc, 95)), "mean_ms": float(np.mean(tc)),
                             "success": okc / 200, "queries_per_s": 1e3 / float(np.mean(tc))}
    # validate a sample of the returned paths against the fp64 oracle
    bad = 0
    for k in np.nonzero(plen > 0)[0][:200]:
        p = paths[k, : plen[k]].astype(np.float64)
        m = ora64.edge_margin(p[:-1], p[1:], snap.as_oracle_scene(), n_steps=0)
        bad += int((m < -1e-4).any())
    res["oracle_invalid_paths_in_200"] = bad
    print(json.dumps(res))

if 6 in which:
    # next-row component 8f-2: batched collision-aware IK (not a BASELINE config)
    snap = sc.goal3_tower()
    pv.set_scene(snap)
    rng = np.random.default_rng(6)
    n = 4096
    pos = np.stack([rng.uniform(0.3, 0.7, n), rng.uniform(-0.4, 0.4, n), rng.uniform(0.15, 0.6, n)], axis=1)
    quat = np.tile(np.array([0.0, 1.0, 0.0, 0.0]), (n, 1))
    res = {"config": "ik", "scene": "goal3_tower", "n_targets": n, "runs": {}}
    for seeds in (32, 128):
        pv.ik_batch(pos[:64], quat[:64], pm.Q_SAFE_HOME, n_seeds=seeds)
        t = time.perf_counter()
        q, ok, err = pv.ik_batch(pos, quat, pm.Q_SAFE_HOME, n_seeds=seeds, seed=3)
        dt = time.perf_counter() - t
        res["runs"][f"seeds_{seeds}"] = {"wall_ms": dt * 1e3, "targets_per_s": n / dt, "solved": float(ok.mean()),
                                         "pos_err_p95_m": float(np.percentile(err[ok, 0], 95)),
                                         "rot_err_p95_rad": float(np.percentile(err[ok, 1], 95))}
    lat = []
    for k in range(100):
        t = time.perf_counter()
        pv.ik_batch(pos[k:k + 1], quat[k:k + 1], pm.Q_SAFE_HOME, n_seeds=128)
        lat.append((time.perf_counter() - t) * 1e3)
    res["single_target_ms"] = {"p50": float(np.median(lat)), "p95": float(np.percentile(lat, 95))}
    print(json.dumps(res))
