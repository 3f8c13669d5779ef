"""Sharding of independent validity work across the GPUs of one box.

Configurations, edges and RRT queries are independent units, so ranks take contiguous shards and the data
path has NO collective.  torch.distributed (NCCL over NVLink on GPUs, gloo in the CPU tests) is used only to
gather the packed verdict words (1 bit per config: 100 M configs -> 12.5 MB in total) and, for a tree shared
across ranks, the per-rank nearest-neighbour candidates (8 B per query per rank).  No float reduction is
involved anywhere, so results are bit-identical for every world size.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(n_total: int, rank: int, world: int, align: int = 32) -> Tuple[int, int]:
    """Contiguous [first, first+count) of rank `rank`; every shard start is a multiple of `align` so the
    packed verdict words of the shards concatenate without bit shifting."""
    units = (n_total + align - 1) // align
    per = (units + world - 1) // world
    first = min(rank * per * align, n_total)
    last = min((rank + 1) * per * align, n_total)
    return first, last - first


def words_per_shard(n_total: int, world: int, align: int = 32) -> int:
    units = (n_total + align - 1) // align
    per = (units + world - 1) // world
    return per * align // 32


def gather_verdict_words(local_words: torch.Tensor, n_total: int, group=None) -> torch.Tensor:
    """All-gather the per-rank verdict words into the full bitmask (int32 words, ceil(n_total/32) long).
    `local_words` holds the words of this rank's shard_range; shorter (last) shards are zero-padded."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    total_words = (n_total + 31) // 32
    if world == 1:
        return local_words[:total_words]
    wps = words_per_shard(n_total, world)
    send = local_words
    if send.numel() != wps:
        send = torch.zeros(wps, dtype=local_words.dtype, device=local_words.device)
        send[: local_words.numel()] = local_words
    out = torch.empty(wps * world, dtype=local_words.dtype, device=local_words.device)
    dist.all_gather_into_tensor(out, send.contiguous(), group=group)
    return out[:total_words]


def sweep_sharded(pv, n_total: int, seed: int, fingers_open: bool = True, group=None):
    """BASELINE config 5: every rank generates and checks its shard of the counter-based config stream on
    its own GPU (no H2D of configs), then the verdict words are all-gathered.  Returns (words, n_valid)."""
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    first, count = shard_range(n_total, rank, world)
    if count > 0:
        words, n_valid = pv.sweep(first, count, seed, fingers_open=fingers_open)
    else:
        words = torch.zeros(0, dtype=torch.int32, device=pv.device)
        n_valid = torch.zeros(1, dtype=torch.int64, device=pv.device)
    full = gather_verdict_words(words, n_total, group)
    if world > 1:
        dist.all_reduce(n_valid, op=dist.ReduceOp.SUM, group=group)  # integer count: exact for any world size
    return full, n_valid


def check_edges_sharded(pv, qa, qb, n_steps: int = 0, resolution: float = 0.0, group=None,
                        gather: Optional["FusedVerdictGather"] = None) -> torch.Tensor:
    """BASELINE config 3 on N GPUs: motions are independent units, so rank r validates the contiguous shard
    shard_range(n, r, world) of the batch (qa, qb: the same (n, 9) tensors, or SoA plane tuples, on every rank) with
    pv.check_edges, and the verdict words (1 bit per edge) are all-gathered.  Every rank returns the words one GPU
    computes for the whole batch.  With `gather` (a FusedVerdictGather of words_per_shard(n, world) words per rank) the
    words travel inside the edge kernel -- peer / multicast stores from its epilogue -- and the only synchronisation is
    the symmetric-memory barrier; the returned mask is then a view of the symmetric buffer (valid until the next use)."""
    from . import panda_model as pm
    if gather is not None:
        n_ = int(qa[0].shape[0] if isinstance(qa, (tuple, list)) else qa.shape[0])
        if gather.words_per_rank != words_per_shard(n_, gather.world):
            raise ValueError("gather slot size does not match words_per_shard(n, world)")
        first, count = shard_range(n_, gather.rank, gather.world)
        if count <= 16384:
            raise ValueError("the fused edge gather applies to shards above 16 384 motions (whole verdict words per warp)")
        cut_ = (lambda t: tuple(x[first:first + count] for x in t)) if isinstance(qa, (tuple, list)) else \
            (lambda t: t[first:first + count])
        gather.hdl.barrier()  # every rank has finished reading the previous mask
        used = (count + 31) // 32
        if used < gather.words_per_rank:
            gather.zero_own_slot_tail(used)
        gather.activate()
        try:
            pv.check_edges(cut_(qa), cut_(qb), n_steps=n_steps,
                           resolution=resolution if resolution > 0 else pm.VALIDITY_RESOLUTION)
        finally:
            gather.deactivate()
        return gather.finish()[: (n_ + 31) // 32]
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    planes = isinstance(qa, (tuple, list))
    n = int(qa[0].shape[0] if planes else qa.shape[0])
    first, count = shard_range(n, rank, world)
    cut = (lambda t: tuple(x[first:first + count] for x in t)) if planes else (lambda t: t[first:first + count])
    if count > 0:
        words = pv.check_edges(cut(qa), cut(qb), n_steps=n_steps,
                               resolution=resolution if resolution > 0 else pm.VALIDITY_RESOLUTION)
    else:
        dev = getattr(pv, "device", None) or torch.device("cpu")
        words = torch.zeros(0, dtype=torch.int32, device=dev)
    return gather_verdict_words(words, n, group)


def sweep_sharded_fused(pv, gather: "FusedVerdictGather", n_total: int, seed: int, fingers_open: bool = True):
    """Same result as sweep_sharded, but the verdict words travel inside the sweep kernel (peer / multicast stores)
    and the only synchronisation is the symmetric-memory barrier.  `gather` must have words_per_rank =
    words_per_shard(n_total, world) and be the handle's active gather (gather.activate()).
    The returned mask is a VIEW of the symmetric buffer, valid until the next sweep is launched into it: an entry
    barrier keeps a faster rank from overwriting words a slower rank is still reading from the previous call."""
    if gather.words_per_rank != words_per_shard(n_total, gather.world):
        raise ValueError(f"gather slot of {gather.words_per_rank} words, but shards of {n_total} configurations over "
                         f"{gather.world} ranks need {words_per_shard(n_total, gather.world)}")
    gather.hdl.barrier()  # every rank has finished reading the previous call's mask
    first, count = shard_range(n_total, gather.rank, gather.world)
    n_valid = torch.zeros(1, dtype=torch.int64, device=pv.device)
    used = (count + 31) // 32
    if used < gather.words_per_rank:
        # a short (or empty) last shard: the tail of this rank's slot must not keep words of an earlier, longer call
        gather.zero_own_slot_tail(used)
    if count > 0:
        _, n_valid = pv.sweep(first, count, seed, fingers_open=fingers_open)
    full = gather.finish()[: (n_total + 31) // 32]
    if gather.world > 1:
        dist.all_reduce(n_valid, op=dist.ReduceOp.SUM, group=gather.group)
    return full, n_valid


def rrtc_batch_sharded(pv, starts, goals, group=None, packed=False, **kw):
    """BASELINE config 4 on N GPUs: the (start, goal) queries are independent, so rank r plans the contiguous shard
    shard_range(n, r, world, align=1) on its own GPU -- with `query_offset` = the shard's first id, which keys the
    random streams by GLOBAL query id -- and the results (path lengths, iteration / check counts, paths) are
    all-gathered.  Every rank returns what one GPU returns for the whole batch: (paths, lengths, iters, checks) as
    numpy arrays.  `pv` needs rrtc_batch(starts, goals, query_offset=..., **kw) -> the same four arrays.
    packed=True returns (states, lengths, iters, checks) with states = the used rows of all paths back to back
    ((sum(lengths), 9); path k starts at lengths[:k].sum()) instead of the dense (n, max_path, 9) block, whose
    allocation dominates the wall time of large batches (4.6 KB per query for typically 2..4 states)."""
    import numpy as np
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    starts = np.ascontiguousarray(starts, dtype=np.float32).reshape(-1, 9)
    goals = np.ascontiguousarray(goals, dtype=np.float32).reshape(-1, 9)
    n = starts.shape[0]
    first, count = shard_range(n, rank, world, align=1)
    paths, plen, iters, checks = pv.rrtc_batch(starts[first:first + count], goals[first:first + count],
                                               query_offset=first, **kw)
    if world == 1:
        if packed:
            return paths[np.arange(paths.shape[1])[None, :] < plen.reshape(-1, 1)], plen, iters, checks
        return paths, plen, iters, checks
    per = (n + world - 1) // world  # shard_range with align=1: every shard but the last holds `per` queries
    max_path = paths.shape[1]
    dev = getattr(pv, "device", None)
    dev = dev if (dev is not None and dist.get_backend(group) == "nccl") else torch.device("cpu")
    # 1. per-query scalars (length, iterations, checks), shards padded to `per` rows
    meta = torch.zeros((per, 3), dtype=torch.int64)
    if count:
        meta[:count] = torch.from_numpy(np.stack([plen, iters, checks], axis=1).astype(np.int64))
    meta_all = torch.empty((world * per, 3), dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(meta_all, meta.to(dev), group=group)
    meta_all = meta_all.cpu().numpy()[:n]
    plen_all = meta_all[:, 0]
    # 2. the path states, packed: only the plen[k] used rows of each path travel (a path buffer is max_path x 9 floats,
    # a typical path 2..4 states), every rank padded to the longest rank's total, which the lengths of step 1 give
    totals = [int(plen_all[min(r * per, n):min((r + 1) * per, n)].sum()) for r in range(world)]
    cap = max(max(totals), 1)
    used = np.arange(max_path)[None, :] < plen.reshape(-1, 1)
    send = torch.zeros((cap, 9), dtype=torch.float32)
    if totals[rank]:
        send[: totals[rank]] = torch.from_numpy(np.ascontiguousarray(paths[used]))
    recv = torch.empty((world * cap, 9), dtype=torch.float32, device=dev)
    dist.all_gather_into_tensor(recv, send.to(dev), group=group)
    recv = recv.cpu().numpy().reshape(world, cap, 9)
    states = np.concatenate([recv[r, : totals[r]] for r in range(world)])
    if packed:
        return states, plen_all.astype(np.int32), meta_all[:, 1].astype(np.int32), meta_all[:, 2].copy()
    allp = np.zeros((n, max_path, 9), dtype=np.float32)
    allp[np.arange(max_path)[None, :] < plen_all[:, None]] = states
    return allp, plen_all.astype(np.int32), meta_all[:, 1].astype(np.int32), meta_all[:, 2].copy()


def merge_nn_candidates(local_d2: torch.Tensor, local_idx: torch.Tensor, group=None):
    """Nearest-tree-node search over a tree sharded across ranks: each rank contributes, per query, the best
    (squared distance, local node index) of its shard; returns (best_d2, owner_rank, owner_local_idx) per query.
    Ties go to the lowest rank, so the winner does not depend on arrival order."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return local_d2, torch.zeros_like(local_idx), local_idx
    b = local_d2.numel()
    d_all = torch.empty(world * b, dtype=local_d2.dtype, device=local_d2.device)
    i_all = torch.empty(world * b, dtype=local_idx.dtype, device=local_idx.device)
    dist.all_gather_into_tensor(d_all, local_d2.contiguous(), group=group)
    dist.all_gather_into_tensor(i_all, local_idx.contiguous(), group=group)
    d_all, i_all = d_all.view(world, b), i_all.view(world, b)
    best = d_all.min(dim=0)
    # torch.min returns the first minimal index on CPU and CUDA for exact ties along dim 0 -> lowest rank
    owner = (d_all == best.values[None]).to(torch.int64).argmax(dim=0)
    return best.values, owner, i_all.gather(0, owner[None]).squeeze(0)


class FusedVerdictGather:
    """Verdict all-gather fused into the validity kernels over NVLink peer memory (no collective launch).

    A symmetric-memory buffer of world x words_per_rank int32 words is allocated on every rank
    (torch.distributed._symmetric_memory, i.e. CUDA VMM handles exchanged once at rendezvous); the handle is told
    (pv_set_gather) to store each verdict word of this rank into every rank's copy at this rank's slot -- through
    the NVSwitch multicast mapping when the fabric offers one, else with one peer store per rank.  `finish()` is
    the only synchronisation: a symmetric-memory barrier after which `self.buf` holds the full mask everywhere.
    """

    def __init__(self, pv, words_per_rank: int, group=None, use_multicast: bool = True):
        import torch.distributed._symmetric_memory as symm
        self.pv = pv
        self.group = group if group is not None else dist.group.WORLD
        self.rank = dist.get_rank(self.group)
        self.world = dist.get_world_size(self.group)
        self.words_per_rank = int(words_per_rank)
        self.buf = symm.empty(self.world * self.words_per_rank, dtype=torch.int32, device=pv.device)
        self.hdl = symm.rendezvous(self.buf, self.group)
        self.buf.zero_()
        mc = 0
        if use_multicast and getattr(self.hdl, "has_multicast_support", False):
            mc = int(self.hdl.multicast_ptr or 0)
        self._mc = mc
        self.multicast = bool(mc)
        self.activate()
        self.hdl.barrier()

    def activate(self, word_offset: int = 0):
        """Make this buffer the handle's gather target (a handle forwards to one buffer at a time).  `word_offset`
        places the NEXT launch's words inside this rank's slot: a caller that gathers several launches into one slot
        (bench.py: one step = 48 launches) moves it before each launch."""
        if not (0 <= word_offset <= self.words_per_rank):
            raise ValueError("word_offset outside this rank's slot")
        self.pv.set_gather(int(self.hdl.buffer_ptrs_dev), self.world, self._mc,
                           self.rank * self.words_per_rank + int(word_offset), self.words_per_rank - int(word_offset))

    def deactivate(self):
        self.pv.set_gather(0, 0, 0, 0, 0)

    def zero_own_slot_tail(self, first_word: int):
        """Zero words [first_word, words_per_rank) of THIS rank's slot on every rank (peer stores; NCCL is not involved)."""
        lo = self.rank * self.words_per_rank + int(first_word)
        hi = (self.rank + 1) * self.words_per_rank
        if hi <= lo:
            return
        for r in range(self.world):
            self.hdl.get_buffer(r, (self.world * self.words_per_rank,), torch.int32)[lo:hi].zero_()

    def finish(self) -> torch.Tensor:
        """All ranks' kernels issued so far have completed and their words are visible: returns the gathered mask."""
        self.hdl.barrier()
        return self.buf

    def close(self):
        self.deactivate()


class FusedCandidateGather:
    """All-gather of the nearest-node candidate records fused into the kernel that finds them (pv_nn_candidates_gather):
    a symmetric-memory buffer of world x capacity x 11 32-bit words on every rank; each rank's kernel stores its records
    into every rank's copy (NVSwitch multicast when the fabric offers it, else peer stores), and `finish(n)` -- a
    symmetric-memory barrier -- returns the (world, n, 11) float view pv_rrtc_steer reads.  No NCCL launch is involved."""

    def __init__(self, pv, capacity: int, group=None, use_multicast: bool = True):
        import torch.distributed._symmetric_memory as symm
        self.pv = pv
        self.group = group if group is not None else dist.group.WORLD
        self.rank = dist.get_rank(self.group)
        self.world = dist.get_world_size(self.group)
        self.capacity = int(capacity)
        self.buf = symm.empty(self.world * self.capacity * 11, dtype=torch.float32, device=pv.device)
        self.hdl = symm.rendezvous(self.buf, self.group)
        self.buf.zero_()
        mc = 0
        if use_multicast and getattr(self.hdl, "has_multicast_support", False):
            mc = int(self.hdl.multicast_ptr or 0)
        self._mc = mc
        self.multicast = bool(mc)
        self.hdl.barrier()

    def launch(self, trees, local_sizes, tree_of, targets):
        if int(targets.shape[0]) > self.capacity:
            raise ValueError("more (tree, target) pairs than the candidate buffer holds")
        self.pv.nn_candidates_gather(trees, local_sizes, tree_of, targets, self.rank, self.world,
                                     int(self.hdl.buffer_ptrs_dev), self._mc)

    def finish(self, n: int) -> torch.Tensor:
        self.hdl.barrier()  # every rank's kernel has completed and its stores are visible
        return self.buf[: self.world * n * 11].view(self.world, n, 11)


class ShardedTreePlanner:
    """Batched multi-query RRT-Connect front end whose TREES are sharded over the GPUs of one box (SURVEY.md 8e,
    BASELINE north_star: "NCCL ... only to gather verdicts and nearest-tree candidates for a batched multi-query
    RRT-Connect front end").

    og.RRTConnect as planning.py:151-156 configures it, for a batch of (start, goal) queries advanced together.  Global
    node g of a tree lives on rank g % world at slot g // world, so the node memory and the nearest-neighbour scan of a
    tree are spread over all ranks.  One round, for the queries still running:

      1. every rank: the nearest of ITS nodes to the round's target (the sample of an EXTEND, the new node of a
         CONNECT) -- pv_nn_candidates, one record of 44 B per query;
      2. all-gather of the records (NCCL);
      3. every rank: reduce by (distance, global index) and steer -- pv_rrtc_steer -- the same motion everywhere;
      4. the motions are an independent batch: rank r validates the contiguous shard shard_range(n, r, world) with
         pv_check_edges (OMPL's DiscreteMotionValidator) and the verdict words are all-gathered (NCCL);
      5. the transitions of RRT-Connect (add the node on its owner rank, swap trees when trapped, ...) as tensor
         operations on replicated control state.

    Samples come from the stream of the single-GPU planner (keyed by seed, GLOBAL query id, iteration), nearest
    neighbours and steering use its expressions, so a query takes the decisions pv_rrtc_batch(replicas=1,
    shortcut_passes=0) takes for it, whatever the world size (the batched edge kernel evaluates sin/cos in hardware, the
    in-kernel validator with the Cody-Waite form: a state within ~1e-6 m of contact could be judged differently).

    `pv` supplies the device steps: nn_candidates, rrtc_steer, rrtc_samples, check_edges, check_states and `.device`
    (PandaValidity; the CPU tests pass a stand-in and run this class over gloo).
    """

    EXTEND, CONNECT, DONE = 0, 1, 4
    SOLVED, ITERCAP, NODECAP, PATHCAP, BADEND = 1, 2, 3, 4, 16

    def __init__(self, pv, max_nodes: int = 2048, group=None, fused_candidates: bool = True):
        self.pv = pv
        self.group = group
        self.fused_candidates = bool(fused_candidates)
        self._cand_gather = None
        self.candidate_exchange = "none"
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.max_nodes = int(max_nodes)
        self.slots = (self.max_nodes + self.world - 1) // self.world
        dev = getattr(pv, "device", None)
        self.dev = dev if dev is not None else torch.device("cpu")
        self.rounds = 0
        self.bytes_gathered = 0

    # -- collectives ------------------------------------------------------------------------------------------------
    def _all_gather(self, t: torch.Tensor) -> torch.Tensor:
        if self.world == 1:
            return t[None]
        out = torch.empty((self.world * t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(out, t.contiguous(), group=self.group)
        self.bytes_gathered += out.numel() * out.element_size()
        return out.view((self.world,) + tuple(t.shape))

    def _valid_sharded(self, ea: torch.Tensor, eb: torch.Tensor, resolution: float) -> torch.Tensor:
        n = ea.shape[0]
        first, count = shard_range(n, self.rank, self.world)
        if count > 0:
            words = self.pv.check_edges(ea[first:first + count], eb[first:first + count], n_steps=0, resolution=resolution)
        else:
            words = torch.zeros(0, dtype=torch.int32, device=self.dev)
        full = gather_verdict_words(words, n, self.group)
        if self.world > 1:
            self.bytes_gathered += words_per_shard(n, self.world) * self.world * 4
        i = torch.arange(n, device=self.dev)
        return ((full[i >> 5] >> (i & 31)) & 1).bool()

    # -- the planner ------------------------------------------------------------------------------------------------
    def solve(self, starts, goals, max_iters: int = 2000, max_path: int = 128, seed: int = 1, rrt_range: float = 0.0,
              resolution: float = 0.0, check_endpoints: bool = False, query_offset: int = 0):
        """Returns (paths: list of (len, 9) float32 arrays -- empty when unsolved --, iters (n,) int32, status (n,) int32),
        identical on every rank."""
        import numpy as np
        from . import panda_model as pm
        rrt_range = float(rrt_range) if rrt_range > 0 else float(pm.RRTC_RANGE)
        resolution = float(resolution) if resolution > 0 else float(pm.VALIDITY_RESOLUTION)
        dev, rank, world, M, slots = self.dev, self.rank, self.world, self.max_nodes, self.slots
        qs = torch.as_tensor(np.ascontiguousarray(starts, dtype=np.float32).reshape(-1, 9)).to(dev)
        qg = torch.as_tensor(np.ascontiguousarray(goals, dtype=np.float32).reshape(-1, 9)).to(dev)
        n = qs.shape[0]
        i32 = dict(dtype=torch.int32, device=dev)
        # (row 2n of `trees` / `parent` is a dump for the writes of rows that have nothing to store, see the transitions)
        trees = torch.zeros((2 * n + 1, 9, slots), dtype=torch.float32, device=dev)  # this rank's slots of every tree
        gsize = torch.ones(2 * n, **i32)                                             # GLOBAL tree sizes (replicated)
        parent = torch.full((2 * n + 1, M), -1, **i32)                               # replicated: 4 B per node
        dump = torch.tensor(2 * n, dtype=torch.long, device=dev)
        zero_l = torch.tensor(0, dtype=torch.long, device=dev)
        c_extend, c_connect, c_done = (torch.tensor(x, **i32) for x in (self.EXTEND, self.CONNECT, self.DONE))
        c_solved, c_nodecap = (torch.tensor(x, **i32) for x in (self.SOLVED, self.NODECAP))
        if rank == 0:  # global node 0 of every tree: rank 0, slot 0
            trees[0:2 * n:2, :, 0] = qs
            trees[1:2 * n:2, :, 0] = qg
        it = torch.zeros(n, **i32)
        cur = torch.zeros(n, **i32)
        phase = torch.full((n,), self.EXTEND, **i32)
        status = torch.zeros(n, **i32)
        target = torch.zeros((n, 9), dtype=torch.float32, device=dev)
        added = torch.zeros(n, **i32)
        conn = torch.full((n,), -1, **i32)
        gsearch = (torch.arange(n, device=dev, dtype=torch.int64) + int(query_offset)).to(torch.int32)
        qid = torch.arange(n, device=dev)

        # candidate exchange: fused into pv_nn_candidates over NVLink peer memory when symmetric memory is available (GPUs,
        # world > 1), else an NCCL / gloo all-gather of the records.  (A round's later verdict all-gather orders round k's
        # readers before round k + 1's writers, so one buffer suffices.)
        fused = None
        if self.fused_candidates and world > 1 and dev.type == "cuda" and hasattr(self.pv, "nn_candidates_gather"):
            try:
                if self._cand_gather is None or self._cand_gather.capacity < n:
                    self._cand_gather = FusedCandidateGather(self.pv, max(n, 1), self.group)
                fused = self._cand_gather
            except Exception:  # no symmetric memory on this fabric: keep the collective
                fused = None
        self.candidate_exchange = ("none" if world == 1 else
                                   ("fused_multicast" if fused is not None and fused.multicast else
                                    "fused_peer_stores" if fused is not None else "all_gather"))

        if check_endpoints and n:
            # OMPL drops invalid / out-of-bounds start and goal states at intake (planning.py:163-187)
            both = torch.cat([qs, qg])
            first, count = shard_range(2 * n, rank, world)
            w = self.pv.check_states(both[first:first + count]) if count else torch.zeros(0, **i32)
            full = gather_verdict_words(w, 2 * n, self.group)
            k = torch.arange(2 * n, device=dev)
            ok = ((full[k >> 5] >> (k & 31)) & 1).bool()
            code = (~ok[:n]).to(torch.int32) + 2 * (~ok[n:]).to(torch.int32)
            status = torch.where(code > 0, self.BADEND + code, status)
            phase = torch.where(code > 0, torch.full_like(phase, self.DONE), phase)

        while True:
            # caps are tested when an EXTEND begins, iteration cap first (pv_rrtc_kernel)
            ext_all = phase == self.EXTEND
            itcap = ext_all & (it >= max_iters)
            ndcap = ext_all & ~itcap & ((gsize[0::2] >= M - 1) | (gsize[1::2] >= M - 1))
            status = torch.where(itcap, torch.full_like(status, self.ITERCAP), status)
            status = torch.where(ndcap, torch.full_like(status, self.NODECAP), status)
            phase = torch.where(itcap | ndcap, torch.full_like(phase, self.DONE), phase)
            a = (phase != self.DONE).nonzero().flatten()
            if a.numel() == 0:
                break
            self.rounds += 1
            ext = phase[a] == self.EXTEND
            tsel = torch.where(ext, cur[a], cur[a] ^ 1)
            tree_of = (2 * a).to(torch.int32) + tsel
            smp = self.pv.rrtc_samples(seed, gsearch[a].contiguous(), it[a].contiguous())
            aim_goal = ext & (it[a] == 0)  # the first extension aims at the goal itself
            tg = torch.where(ext[:, None], torch.where(aim_goal[:, None], qg[a], smp), target[a]).contiguous()
            local_sizes = torch.clamp((gsize - rank + world - 1) // world, min=0).to(torch.int32)
            if fused is not None:
                fused.launch(trees, local_sizes, tree_of.contiguous(), tg)
                all_cand = fused.finish(int(a.numel()))
                self.bytes_gathered += all_cand.numel() * 4
            else:
                all_cand = self._all_gather(self.pv.nn_candidates(trees, local_sizes, tree_of.contiguous(), tg, rank, world))
            from_g, ea, eb, reach = self.pv.rrtc_steer(all_cand, tg, rrt_range)
            valid = self._valid_sharded(ea, eb, resolution)
            reach = reach.bool()

            # ---- transitions of RRT-Connect, as whole-array selects on the active rows (no boolean-mask indexing: every
            # masked gather / scatter would be a device synchronisation) ---------------------------------------------------
            to = tree_of.long()
            ni = gsize[to]                       # global index the new node would get
            v = valid
            ext_v, con_v = v & ext, v & ~ext
            # the new node is stored by its owner rank only; everybody else (and invalid motions) writes to a dump tree
            own = v & ((ni % world) == rank)
            trees[torch.where(own, to, dump), :, torch.where(own, (ni // world).long(), zero_l)] = eb
            parent[torch.where(v, to, dump), torch.where(v, ni.long(), zero_l)] = from_g
            gsize[to] = ni + v.to(torch.int32)   # (one tree per active query: no duplicate indices)
            target[a] = torch.where(ext_v[:, None], eb, target[a])
            added[a] = torch.where(ext_v, ni, added[a])
            solved = con_v & reach
            conn[a] = torch.where(solved, ni, conn[a])
            full_ = con_v & ~reach & (ni + 1 >= M - 1)
            ph = phase[a]
            ph = torch.where(ext_v, c_connect, ph)
            ph = torch.where(solved | full_, c_done, ph)
            ph = torch.where(~v, c_extend, ph)       # trapped: next iteration, the trees swap roles
            phase[a] = ph
            st_a = status[a]
            st_a = torch.where(solved, c_solved, st_a)
            status[a] = torch.where(full_, c_nodecap, st_a)
            cur_a = cur[a]
            cur[a] = torch.where(v, cur_a, cur_a ^ 1)
            it[a] = it[a] + (~v).to(torch.int32)

        # ---- path extraction: index chains from the replicated parents, states from their owner ranks -----------------
        st = status.cpu().numpy()
        iters = torch.where(status == self.SOLVED, it + 1, it).cpu().numpy().astype(np.int32)
        solved = np.nonzero(st == self.SOLVED)[0]
        par = parent.cpu().numpy()
        cur_h, add_h, con_h = cur.cpu().numpy(), added.cpu().numpy(), conn.cpu().numpy()
        chains, want_t, want_g = {}, [], []
        for q in solved:
            i_s, i_g = (add_h[q], con_h[q]) if cur_h[q] == 0 else (con_h[q], add_h[q])
            s_chain = []
            x = int(i_s)
            while x >= 0:
                s_chain.append(x)
                x = int(par[2 * q, x])
            g_chain = []
            x = int(par[2 * q + 1, i_g])  # the goal tree's copy of the connecting state is skipped
            while x >= 0:
                g_chain.append(x)
                x = int(par[2 * q + 1, x])
            chain = [(2 * q, g) for g in reversed(s_chain)] + [(2 * q + 1, g) for g in g_chain]
            if len(chain) > max_path:
                st[q] = self.PATHCAP
                continue
            chains[int(q)] = (len(want_t), len(chain))
            want_t += [c[0] for c in chain]
            want_g += [c[1] for c in chain]
        paths = [np.zeros((0, 9), np.float32) for _ in range(n)]
        if want_t:
            wt = torch.as_tensor(want_t, device=dev, dtype=torch.long)
            wg = torch.as_tensor(want_g, device=dev, dtype=torch.long)
            mine = (wg % world) == rank
            rows = torch.zeros((len(want_t), 9), dtype=torch.float32, device=dev)
            rows[mine] = trees[wt[mine], :, wg[mine] // world]
            allr = self._all_gather(rows)                               # (world, total, 9)
            states = allr[wg % world, torch.arange(len(want_t), device=dev)].cpu().numpy()
            for q, (off, ln) in chains.items():
                paths[q] = states[off:off + ln].copy()
        return paths, iters, st.astype(np.int32)
