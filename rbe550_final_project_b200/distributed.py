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


def sweep_sharded_fused(pv, gather: "FusedVerdictGather", n_total: int, seed: int, fingers_open: bool = True):
    """Same result as sweep_sharded, but the verdict words travel inside the sweep kernel (peer / multicast stores)
    and the only synchronisation is the symmetric-memory barrier.  `gather` must have words_per_rank =
    words_per_shard(n_total, world)."""
    first, count = shard_range(n_total, gather.rank, gather.world)
    n_valid = torch.zeros(1, dtype=torch.int64, device=pv.device)
    if count > 0:
        _, n_valid = pv.sweep(first, count, seed, fingers_open=fingers_open)
    full = gather.finish()[: (n_total + 31) // 32]
    if gather.world > 1:
        dist.all_reduce(n_valid, op=dist.ReduceOp.SUM, group=gather.group)
    return full, n_valid


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
        self.multicast = bool(mc)
        pv.set_gather(int(self.hdl.buffer_ptrs_dev), self.world, mc, self.rank * self.words_per_rank, self.words_per_rank)
        self.hdl.barrier()

    def finish(self) -> torch.Tensor:
        """All ranks' kernels issued so far have completed and their words are visible: returns the gathered mask."""
        self.hdl.barrier()
        return self.buf

    def close(self):
        self.pv.set_gather(0, 0, 0, 0, 0)
