"""world_size-2 gloo tests (CPU) of the sharding / gather logic used for the multi-GPU sweeps."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n_total, ret):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import panda_oracle as po
    from rbe550_final_project_b200 import panda_model as pm, scenes as sc
    from rbe550_final_project_b200.distributed import gather_verdict_words, merge_nn_candidates, shard_range
    model = pm.model_arrays()
    scene = sc.goal1_scattered().as_oracle_scene()
    first, count = shard_range(n_total, rank, world)
    q = po.sweep_configs(first, count, 99, model)
    valid = po.state_margin(q.astype(np.float64), scene, model) >= 0
    words = torch.from_numpy(po.pack_bits(valid).view(np.int32).copy())
    full = gather_verdict_words(words, n_total)
    # sharded nearest neighbour: rank r owns tree nodes r, r+world, ...
    rng = np.random.default_rng(5)
    tree = rng.uniform(-1, 1, size=(101, 9))
    queries = rng.uniform(-1, 1, size=(17, 9))
    mine = tree[rank::world]
    d2 = ((queries[:, None, :] - mine[None]) ** 2).sum(-1)
    li = d2.argmin(1)
    bd, owner, idx = merge_nn_candidates(torch.from_numpy(d2.min(1)), torch.from_numpy(li))
    if rank == 0:
        ret["full"] = full.numpy().copy()
        ret["nn"] = (owner.numpy() + world * idx.numpy()).copy()
        ret["nn_ref"] = ((queries[:, None, :] - tree[None]) ** 2).sum(-1).argmin(1)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_total", [4096, 5000])
def test_sharded_gather_matches_single_rank(n_total):
    sys.path.insert(0, ROOT)
    from oracle import panda_oracle as po
    from rbe550_final_project_b200 import panda_model as pm, scenes as sc
    model = pm.model_arrays()
    q = po.sweep_configs(0, n_total, 99, model)
    valid = po.state_margin(q.astype(np.float64), sc.goal1_scattered().as_oracle_scene(), model) >= 0
    ref = po.pack_bits(valid)
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, n_total, ret), nprocs=2, join=True)
    assert np.array_equal(ret["full"].view(np.uint32), ref)
    assert np.array_equal(ret["nn"], ret["nn_ref"])


def test_shard_range_properties():
    from rbe550_final_project_b200.distributed import shard_range, words_per_shard
    for n in (1, 31, 32, 33, 1000, 104857600):
        for world in (1, 2, 4, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert sum(c for _, c in spans) == n
            pos = 0
            for first, count in spans:
                assert first == pos or count == 0
                assert first % 32 == 0 or count == 0
                pos = first + count
            assert words_per_shard(n, world) * world * 32 >= n
