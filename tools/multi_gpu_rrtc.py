"""BASELINE config 4 on N GPUs: batched multi-query RRT-Connect (tall-tower scene), the queries sharded over the ranks of
one box, results all-gathered (distributed.rrtc_batch_sharded).  Launch:
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29513 \
      tools/multi_gpu_rrtc.py [n_queries ...]
Every rank must end up with the rows one GPU computes for the whole batch (random streams keyed by global query id);
rank 0 checks that against its own unsharded run and prints one JSON line per batch size."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from rbe550_final_project_b200 import panda_model as pm, scenes as sc
from rbe550_final_project_b200.distributed import rrtc_batch_sharded
from rbe550_final_project_b200.validity import PandaValidity, unpack_bits

sizes = [int(a) for a in sys.argv[1:]] or [4096, 65536]
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
pv = PandaValidity(local)
pv.set_scene(sc.goal3_tower())
rng = np.random.default_rng(4096)
need = 2 * max(sizes)
cand = rng.uniform(pm.Q_LOWER, pm.Q_UPPER, size=(int(need * 7.5), 9)).astype(np.float32); cand[:, 7:] = 0.04
ok = unpack_bits(pv.check_states_host(cand), len(cand))
poses = pv.fk(torch.as_tensor(cand, device="cuda")).cpu().numpy()
ok &= poses[:, 8, 2] > 0.15  # hand z > 0.15
valid = cand[ok]
assert len(valid) >= need, (len(valid), need)
kw = dict(max_iters=2000, max_nodes=2048, max_path=128, seed=7, replicas=1, shortcut_passes=2)
for nq in sizes:
    starts, goals = valid[:nq], valid[nq:2 * nq]
    rrtc_batch_sharded(pv, starts, goals, **kw)  # warm-up: sizes the device arena and the pinned mirrors
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t = time.perf_counter()
    states, plen, iters, checks = rrtc_batch_sharded(pv, starts, goals, packed=True, **kw)
    torch.cuda.synchronize()
    dt = torch.tensor([time.perf_counter() - t], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    paths = rrtc_batch_sharded(pv, starts, goals, **kw)[0]  # dense form, for the comparison below (not timed)
    # the planning alone (no gather): what each rank spends on its shard
    first = min(rank * ((nq + world - 1) // world), nq)
    cnt = min(first + (nq + world - 1) // world, nq) - first
    t = time.perf_counter()
    pv.rrtc_batch(starts[first:first + cnt], goals[first:first + cnt], query_offset=first, **kw)
    ds = torch.tensor([time.perf_counter() - t], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ds, op=dist.ReduceOp.MAX)
    if rank == 0:
        ref = pv.rrtc_batch(starts, goals, **kw)
        used = np.arange(ref[0].shape[1])[None, :] < ref[1][:, None]  # rows beyond a path's length are scratch
        same = all(np.array_equal(a, b) for a, b in zip((paths[used], plen, iters, checks), (ref[0][used],) + tuple(ref[1:])))
        same = same and np.array_equal(states, ref[0][np.arange(ref[0].shape[1])[None, :] < ref[1][:, None]])
        solved = plen > 0
        print(json.dumps({"config": 4, "scene": "goal3_tower", "n_queries": nq, "n_gpus": world,
                          "wall_ms_incl_gather": float(dt.item()) * 1e3, "queries_per_s": nq / float(dt.item()),
                          "wall_ms_shard_only": float(ds.item()) * 1e3, "queries_per_s_shard_only": nq / float(ds.item()),
                          "success": float(solved.mean()), "iters_p50": float(np.median(iters[solved])),
                          "identical_to_one_gpu": bool(same), "gathered_bytes": int(states.nbytes + 24 * nq),
                          "result": "packed path states + lengths, iterations, checks on every rank"}), flush=True)
if world > 1:
    dist.destroy_process_group()
