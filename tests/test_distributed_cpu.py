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


class _OraclePlanner:
    """Stands in for PandaValidity in the CPU test: the C restatement of the device planner, one search per query,
    random streams keyed by global query id like pv_rrtc_batch(query_offset=...)."""
    device = None

    def __init__(self):
        from oracle.c_oracle import COracle
        from rbe550_final_project_b200 import panda_model as pm, scenes as sc
        self.ora = COracle(pm.model_arrays(), "f32")
        wall = sc.make_obb((0.55, 0.0, 0.35), (0.5, 0.04, 0.7))
        self.scene = sc.SceneSnapshot(obb=np.array([wall], dtype=np.float32), names=["wall"], entity_idx=[1]).as_oracle_scene()

    def rrtc_batch(self, starts, goals, query_offset=0, max_path=64, seed=1, **kw):
        n = len(starts)
        paths = np.zeros((n, max_path, 9), np.float32)
        plen, iters, checks = np.zeros(n, np.int32), np.zeros(n, np.int32), np.zeros(n, np.int64)
        for k in range(n):
            p, it, ch = self.ora.rrtc(starts[k], goals[k], self.scene, seed=seed, search=query_offset + k, max_path=max_path,
                                      max_iters=300, max_nodes=512)
            paths[k, : len(p)] = p
            plen[k], iters[k], checks[k] = len(p), it, ch
        return paths, plen, iters, checks


def _rrtc_queries(n):
    from rbe550_final_project_b200 import panda_model as pm
    rng = np.random.default_rng(8)
    a = rng.uniform(pm.Q_LOWER, pm.Q_UPPER, size=(n, 9)).astype(np.float32)
    b = rng.uniform(pm.Q_LOWER, pm.Q_UPPER, size=(n, 9)).astype(np.float32)
    return a, b


def _rrtc_worker(rank, world, port, n, ret):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from rbe550_final_project_b200.distributed import rrtc_batch_sharded
    a, b = _rrtc_queries(n)
    out = rrtc_batch_sharded(_OraclePlanner(), a, b, max_path=64, seed=5)
    packed = rrtc_batch_sharded(_OraclePlanner(), a, b, packed=True, max_path=64, seed=5)
    used = np.arange(out[0].shape[1])[None, :] < out[1][:, None]
    assert np.array_equal(packed[0], out[0][used]) and all(np.array_equal(x, y) for x, y in zip(packed[1:], out[1:]))
    ret[rank] = [np.asarray(x).copy() for x in out]
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [13, 16])
def test_sharded_rrtc_matches_single_rank(n):
    """world_size 2 over gloo: every rank ends up with the rows one rank computes for the whole batch (ragged last shard
    included), because the random streams follow the global query id."""
    a, b = _rrtc_queries(n)
    ref = _OraclePlanner().rrtc_batch(a, b, max_path=64, seed=5)
    assert (ref[2] > 1).any() and (ref[1] > 0).any()
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_rrtc_worker, args=(2, port, n, ret), nprocs=2, join=True)
    for rank in (0, 1):
        for j in range(4):
            assert np.array_equal(ret[rank][j], ref[j]), (rank, j)


class _FakePlanner:
    """Deterministic stand-in: path k has (global id * 7) % 6 states (0 = unsolved) whose entries encode (id, row)."""
    device = None

    def rrtc_batch(self, starts, goals, query_offset=0, max_path=8, **kw):
        n = len(starts)
        ids = query_offset + np.arange(n)
        plen = ((ids * 7) % 6).astype(np.int32)
        paths = np.full((n, max_path, 9), -1.0, np.float32)  # rows beyond a path's length: scratch, must not travel
        for k in range(n):
            for r in range(plen[k]):
                paths[k, r] = ids[k] * 100 + r + starts[k, 0]
        return paths, plen, (ids % 5).astype(np.int32), (ids * 3).astype(np.int64)


def _fake_worker(rank, world, port, n, ret):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from rbe550_final_project_b200.distributed import rrtc_batch_sharded
    a = np.arange(n * 9, dtype=np.float32).reshape(n, 9) / 1000.0
    dense = rrtc_batch_sharded(_FakePlanner(), a, a, max_path=8)
    packed = rrtc_batch_sharded(_FakePlanner(), a, a, packed=True, max_path=8)
    ret[rank] = ([np.asarray(x).copy() for x in dense], [np.asarray(x).copy() for x in packed])
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n", [(2, 1), (3, 10), (3, 12), (4, 3)])
def test_sharded_rrtc_packing(world, n):
    """Ragged shards (also empty ones: more ranks than queries), unsolved queries and scratch rows: the gathered result
    equals the one-rank result in its used rows, is zero elsewhere, and the packed form is the used rows back to back."""
    a = np.arange(n * 9, dtype=np.float32).reshape(n, 9) / 1000.0
    ref = _FakePlanner().rrtc_batch(a, a, max_path=8)
    used = np.arange(8)[None, :] < ref[1][:, None]
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 33500 + (os.getpid() % 2000) + world
    mp.spawn(_fake_worker, args=(world, port, n, ret), nprocs=world, join=True)
    for rank in range(world):
        dense, packed = ret[rank]
        assert np.array_equal(dense[0][used], ref[0][used]) and (dense[0][~used] == 0).all()
        assert np.array_equal(packed[0], ref[0][used])
        for j in (1, 2, 3):
            assert np.array_equal(dense[j], ref[j]) and np.array_equal(packed[j], ref[j]), (rank, j)
