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


# ---- sharded-TREE planner front end (distributed.ShardedTreePlanner) over gloo ---------------------------------------
def _fma32(a, b, c):
    """fp32 fma: a * b is exact in fp64 (24 + 24 bits), one rounding to fp64 and one to fp32 follow."""
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(np.float32)


class _OracleSteps:
    """CPU stand-in for the device steps PandaValidity gives ShardedTreePlanner (pv_nn_candidates, pv_rrtc_steer,
    pv_rrtc_samples, pv_check_edges, pv_check_states), restated with numpy fp32 and the fp32 C oracle."""
    device = None

    def __init__(self):
        from oracle.c_oracle import COracle
        from rbe550_final_project_b200 import panda_model as pm, scenes as sc
        self.pm = pm
        self.model = pm.model_arrays()
        self.ora = COracle(self.model, "f32")
        wall = sc.make_obb((0.55, 0.0, 0.35), (0.5, 0.04, 0.7))
        self.scene = sc.SceneSnapshot(obb=np.array([wall], dtype=np.float32), names=["wall"], entity_idx=[1]).as_oracle_scene()

    def rrtc_samples(self, seed, gsearch, it):
        from oracle import panda_oracle as po
        gs, itn = gsearch.numpy().astype(np.uint32), it.numpy().astype(np.uint32)
        n = len(gs)
        u = []
        for blk in range(3):
            ctr = np.stack([itn, gs, np.full(n, blk, np.uint32), np.ones(n, np.uint32)], axis=1)
            r = po.philox4x32(ctr, (np.uint32(seed), np.uint32(0x52525443)))
            u.append((r >> np.uint32(8)).astype(np.float32) * np.float32(2.0 ** -24))
        u = np.concatenate(u, axis=1)[:, :9]
        lo, hi = self.model["q_lower"].astype(np.float32), self.model["q_upper"].astype(np.float32)
        return torch.from_numpy(_fma32(u, np.broadcast_to((hi - lo).astype(np.float32), u.shape), np.broadcast_to(lo, u.shape)))

    def nn_candidates(self, trees, sizes, tree_of, targets, rank, world):
        tr, sz, to, tg = trees.numpy(), sizes.numpy(), tree_of.numpy(), targets.numpy()
        out = np.zeros((len(to), 11), np.float32)
        for t, tree in enumerate(to):
            size = int(sz[tree])
            if size == 0:
                out[t, 0] = np.float32(3.0e38)
                out[t, 1:2] = np.array([0x7FFFFFFF], np.int32).view(np.float32)
                continue
            d2 = np.zeros(size, np.float32)
            for k in range(9):
                d = (tr[tree, k, :size] - tg[t, k]).astype(np.float32)
                d2 = _fma32(d, d, d2)
            i = int(np.argmin(d2))  # first minimum = lowest slot = lowest global index
            out[t, 0] = d2[i]
            out[t, 1:2] = np.array([i * world + rank], np.int32).view(np.float32)
            out[t, 2:] = tr[tree, :, i]
        return torch.from_numpy(out)

    def rrtc_steer(self, cand, targets, rrt_range):
        c, tg = cand.numpy(), targets.numpy()
        world, n = c.shape[0], c.shape[1]
        gi_all = np.ascontiguousarray(c[:, :, 1]).view(np.int32)
        from_g, reach = np.zeros(n, np.int32), np.zeros(n, np.int32)
        ea, eb = np.zeros((n, 9), np.float32), np.zeros((n, 9), np.float32)
        rng = np.float32(rrt_range)
        for i in range(n):
            r = min(range(world), key=lambda r: (c[r, i, 0], gi_all[r, i]))
            bd = c[r, i, 0]
            d = np.sqrt(bd, dtype=np.float32)
            ea[i] = c[r, i, 2:]
            from_g[i] = gi_all[r, i]
            if d > rng:
                f = np.float32(rng / d)
                eb[i] = _fma32(np.full(9, f, np.float32), (tg[i] - ea[i]).astype(np.float32), ea[i])
            else:
                eb[i] = tg[i]
                reach[i] = 1
        return torch.from_numpy(from_g), torch.from_numpy(ea), torch.from_numpy(eb), torch.from_numpy(reach)

    def check_edges(self, qa, qb, n_steps=0, resolution=0.0):
        from oracle import panda_oracle as po
        m = self.ora.edge_margin(qa.numpy(), qb.numpy(), self.scene, n_steps=n_steps, resolution=resolution)
        return torch.from_numpy(po.pack_bits(m >= 0).view(np.int32).copy())

    def check_states(self, q):
        from oracle import panda_oracle as po
        m = self.ora.state_margin(q.numpy(), self.scene)
        return torch.from_numpy(po.pack_bits(m >= 0).view(np.int32).copy())


def _tree_worker(rank, world, port, n, max_nodes, ret):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from rbe550_final_project_b200.distributed import ShardedTreePlanner
    a, b = _rrtc_queries(n)
    pl = ShardedTreePlanner(_OracleSteps(), max_nodes=max_nodes)
    paths, iters, status = pl.solve(a, b, max_iters=300, max_path=64, seed=5)
    ret[rank] = ([p.copy() for p in paths], iters.copy(), status.copy(), pl.rounds, pl.bytes_gathered)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_tree_planner_matches_the_single_tree_planner(world):
    """Trees sharded node by node over the ranks of a gloo group (nearest-node candidates and motion verdicts gathered
    every round): every rank returns, query by query, the path and iteration count of the C restatement of the
    single-GPU planner (one search per query, no shortcutting), whose trees live in one memory."""
    n, max_nodes = 13, 512
    a, b = _rrtc_queries(n)
    eng = _OraclePlanner()
    ref = [eng.ora.rrtc(a[k], b[k], eng.scene, seed=5, search=k, max_path=64, max_iters=300, max_nodes=max_nodes,
                        shortcut_passes=0) for k in range(n)]
    assert any(r[1] > 1 for r in ref) and any(len(r[0]) > 2 for r in ref)
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 35500 + (os.getpid() % 2000) + world
    mp.spawn(_tree_worker, args=(world, port, n, max_nodes, ret), nprocs=world, join=True)
    for rank in range(world):
        paths, iters, status, rounds, nbytes = ret[rank]
        assert rounds >= 2 and nbytes > 0
        for k in range(n):
            assert iters[k] == ref[k][1], (rank, k)
            assert np.array_equal(paths[k], ref[k][0]), (rank, k)
            assert (status[k] == 1) == (len(ref[k][0]) > 0)


def _edges_worker(rank, world, port, n, ret):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from rbe550_final_project_b200.distributed import check_edges_sharded
    a, b = _rrtc_queries(n)
    full = check_edges_sharded(_OracleSteps(), torch.from_numpy(a), torch.from_numpy(b), n_steps=0)
    ret[rank] = full.numpy().copy()
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n", [(2, 100), (3, 33), (4, 7)])
def test_sharded_edge_batches_match_single_rank(world, n):
    """BASELINE config 3's sharding: contiguous 32-aligned shards of a motion batch per rank (ragged and empty shards
    included), verdict words all-gathered: every rank holds the words of the unsplit call."""
    a, b = _rrtc_queries(n)
    ref = _OracleSteps().check_edges(torch.from_numpy(a), torch.from_numpy(b), n_steps=0, resolution=0.13037159046356686).numpy()
    assert 0 < np.unpackbits(ref.view(np.uint8)).sum() < n
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 37500 + (os.getpid() % 2000) + world
    mp.spawn(_edges_worker, args=(world, port, n, ret), nprocs=world, join=True)
    for rank in range(world):
        assert np.array_equal(ret[rank], ref), rank
