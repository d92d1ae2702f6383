"""Synthetic graph batches of the shapes BASELINE.json names (no datasets are reachable offline).

A batch is a disjoint union of graphs in the reference's format (train_pytorch_U2GNN_Sup.py:91-119):
input_x[N, k+1] int64 = [node, k neighbours sampled uniformly with replacement inside the node's
graph], rowptr[G+1] (CSR form of graph_pool), X[N, d] fp32 features, labels[G].  Generated on the
device with a seeded torch.Generator; this is bench/test input plumbing, not part of the hot path.
"""
import torch


def make_batch(n_nodes, k, d, num_classes=2, avg_graph=61, seed=2024, device="cuda", vocab=None):
    g = torch.Generator(device=device).manual_seed(seed)
    lo, hi = max(2, avg_graph // 2), avg_graph + avg_graph // 2
    G = max(1, n_nodes // avg_graph)
    sizes = torch.randint(lo, hi + 1, (G,), generator=g, device=device, dtype=torch.int64)
    # rescale so the sizes sum to exactly n_nodes
    diff = int(n_nodes - sizes.sum().item())
    sizes[-1] = max(1, sizes[-1] + diff) if abs(diff) < lo else sizes[-1]
    if int(sizes.sum().item()) != n_nodes:
        per = n_nodes // G
        sizes = torch.full((G,), per, device=device, dtype=torch.int64)
        sizes[: n_nodes - per * G] += 1
    rowptr = torch.zeros(G + 1, dtype=torch.int64, device=device)
    rowptr[1:] = torch.cumsum(sizes, 0)
    gid = torch.repeat_interleave(torch.arange(G, device=device), sizes)
    start, size = rowptr[gid], sizes[gid]
    r = torch.rand((n_nodes, k), generator=g, device=device)
    nbr = start[:, None] + torch.minimum((r * size[:, None]).long(), size[:, None] - 1)
    input_x = torch.cat([torch.arange(n_nodes, device=device)[:, None], nbr], 1).contiguous()
    X = torch.randn((n_nodes, d), generator=g, device=device)
    labels = torch.randint(0, num_classes, (G,), generator=g, device=device, dtype=torch.int64)
    out = dict(input_x=input_x, rowptr=rowptr, X=X, labels=labels, G=G)
    if vocab is not None:
        out["input_y"] = torch.randperm(vocab, generator=g, device=device)[:n_nodes].contiguous()
    return out
