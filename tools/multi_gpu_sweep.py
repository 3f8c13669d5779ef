"""BASELINE config 5: N-config validity sweep sharded over the ranks of one box, verdict words gathered with
NCCL.  Launch: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
--master-port 29511 tools/multi_gpu_sweep.py [n_total].  Rank 0 re-checks a slice on its own GPU and
against the CPU oracle, and prints one JSON line."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from rbe550_final_project_b200 import panda_model as pm, scenes as sc
from rbe550_final_project_b200.distributed import shard_range, sweep_sharded
from rbe550_final_project_b200.validity import PandaValidity, unpack_bits

n_total = int(sys.argv[1]) if len(sys.argv) > 1 else 104_857_600
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
pv = PandaValidity(local)
pv.set_scene(sc.goal1_scattered())
seed = 20251212
for _ in range(2):
    full, n_valid = sweep_sharded(pv, n_total, seed)
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 3
e0.record()
for _ in range(reps):
    full, n_valid = sweep_sharded(pv, n_total, seed)
e1.record()
torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1) / reps], device="cuda")
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
# ---- fused variant: verdict words stored into every rank's buffer from inside the sweep kernel ----------------
fused = None
if world > 1:
    try:
        from rbe550_final_project_b200.distributed import FusedVerdictGather, sweep_sharded_fused, words_per_shard
        g = FusedVerdictGather(pv, words_per_shard(n_total, world))
        for _ in range(2):
            f_full, f_valid = sweep_sharded_fused(pv, g, n_total, seed)
        dist.barrier(); torch.cuda.synchronize()
        e0.record()
        for _ in range(reps):
            f_full, f_valid = sweep_sharded_fused(pv, g, n_total, seed)
        e1.record(); torch.cuda.synchronize()
        fms = torch.tensor([e0.elapsed_time(e1) / reps], device="cuda")
        dist.all_reduce(fms, op=dist.ReduceOp.MAX)
        same_t = torch.tensor([int(torch.equal(f_full, full))], device="cuda")
        dist.all_reduce(same_t, op=dist.ReduceOp.MIN)
        fused = {"ms": float(fms.item()), "checks_per_s": n_total / (float(fms.item()) * 1e-3), "multicast": g.multicast,
                 "identical_to_nccl_gather_on_all_ranks": bool(same_t.item()), "n_valid": int(f_valid.item())}
        g.close()
    except Exception as exc:
        fused = {"error": repr(exc)[:300]}
if rank == 0:
    # invariance: the gathered mask equals what one GPU computes for the same index range
    k = min(n_total, 4_194_304)
    ref_bits, ref_cnt = pv.sweep(0, k, seed)
    same = bool(torch.equal(full[: (k + 31) // 32], ref_bits))
    # and a slice against the CPU oracle (outside the 1e-4 band)
    from oracle import panda_oracle as po
    from oracle.c_oracle import COracle
    m = 200_000
    q = po.sweep_configs(0, m, seed, pm.model_arrays())
    margin = COracle(pm.model_arrays(), "f64").state_margin(q.astype(np.float64), sc.goal1_scattered().as_oracle_scene())
    gpu = unpack_bits(full[: (m + 31) // 32], m)
    far = np.abs(margin) > 1e-4
    ok = bool((gpu[far] == (margin[far] >= 0)).all())
    total_valid = int(n_valid.item())
    popc = int(np.unpackbits(full.cpu().numpy().view(np.uint8)).sum())
    print(json.dumps({"config": "validity sweep, goal1 scene, device-generated configs, verdict all-gather",
                      "n_total": n_total, "n_gpus": world, "ms": float(ms.item()),
                      "checks_per_s": n_total / (float(ms.item()) * 1e-3), "n_valid": total_valid,
                      "count_matches_mask": popc == total_valid, "matches_single_gpu": same, "matches_oracle": ok,
                      "gather_bytes": int(full.numel() * 4), "fused_peer_gather": fused}))
if world > 1:
    dist.destroy_process_group()
