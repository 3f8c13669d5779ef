"""The sharded-TREE planner front end on N GPUs (distributed.ShardedTreePlanner): every tree of a batch of RRT-Connect
queries is dealt node by node to the ranks; per round the ranks exchange nearest-node candidates and motion verdicts over
NCCL.  Launch:
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29514 \
      tools/multi_gpu_tree.py [n_queries]
Every rank checks its result against the one-kernel planner (pv_rrtc_batch, one search per query, no shortcutting) run on
its own GPU; rank 0 prints one JSON line."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import bench
from rbe550_final_project_b200.validity import PandaValidity

nq = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
pv = PandaValidity(local)
rec = bench.bench_tree(torch, dist, pv, rank, world, nq=nq)
rec["n_gpus"] = world
if rank == 0:
    print(json.dumps(rec), flush=True)
if world > 1:
    dist.destroy_process_group()
