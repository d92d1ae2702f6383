"""Data-parallel host logic (one process per GPU, torch.distributed; NCCL on the GPUs, gloo in CPU tests).

The reference is single-process (no collective call site anywhere), so everything here
is new.  The path shards naturally: a batch is a disjoint union of graphs, sampled
neighbours never leave their graph and pooling is per graph, so rank r owns a contiguous range of graphs
(balanced by node count) with batch-local node ids; the model is replicated.  Collectives per step:
  supervised    one all-reduce(sum) of the flat gradient arena (the squared norm is recomputed locally from
                the reduced gradient, so it is identical on every rank)
  unsupervised  the class table [V, D] is row-sharded in contiguous blocks aligned with graph ownership, so
                the true-class rows of a rank's nodes are always local; the `ns` sampled rows are assembled
                on every rank with one all-reduce of an [ns, D] buffer and their gradient goes back the same
                way (ns*D*4 B = 8 KB at ns 512, D 4)
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def balanced_graph_ranges(rowptr: torch.Tensor, world_size: int):
    """Contiguous graph ranges [g0, g1) per rank with (nearly) equal node counts.  rowptr on any device."""
    rp = rowptr.detach().cpu()
    G = rp.numel() - 1
    total = int(rp[-1])
    bounds = [0]
    for r in range(1, world_size):
        target = total * r // world_size
        g = int(torch.searchsorted(rp, torch.tensor(target), right=False))
        g = min(max(g, bounds[-1]), G)
        bounds.append(g)
    bounds.append(G)
    return [(bounds[r], bounds[r + 1]) for r in range(world_size)]


def shard_graph_batch(input_x, rowptr, X, labels, rank, world_size):
    """Local part of a supervised batch for `rank`: node ids are re-based to the shard."""
    g0, g1 = balanced_graph_ranges(rowptr, world_size)[rank]
    n0, n1 = int(rowptr[g0]), int(rowptr[g1])
    return dict(input_x=(input_x[n0:n1] - n0).contiguous(), rowptr=(rowptr[g0:g1 + 1] - n0).contiguous(),
                X=X[n0:n1].contiguous(), labels=labels[g0:g1].contiguous(), node_range=(n0, n1), graph_range=(g0, g1))


def all_reduce_sum_(t: torch.Tensor):
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


class RowShard:
    """Contiguous row sharding of the unsupervised class table: rank r owns rows [lo_r, hi_r)."""

    def __init__(self, vocab_size, world_size, rank, bounds=None):
        if bounds is None:
            bounds = [vocab_size * r // world_size for r in range(world_size + 1)]
        self.bounds, self.rank, self.world = list(bounds), rank, world_size
        self.lo, self.hi = self.bounds[rank], self.bounds[rank + 1]

    @property
    def local_rows(self):
        return self.hi - self.lo

    def owns(self, ids: torch.Tensor):
        return (ids >= self.lo) & (ids < self.hi)

    def to_local(self, ids: torch.Tensor):
        return ids - self.lo
