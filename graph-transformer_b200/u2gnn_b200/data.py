"""Dataset front-end and host batch builder for the CLI / accuracy runs.

Own implementation of what the reference does with networkx in `util.load_data` (U2GNN_pytorch/util.py:54-158)
and `get_batch_data` (train_pytorch_U2GNN_Sup.py:91-119): the text format is parsed straight into per-graph
neighbour lists (numpy), features are the one-hot node tag (or degree, `degree_as_tag`), folds come from the
same `StratifiedKFold(10, shuffle=True, random_state=0)`.  This is plumbing around the hot path (SURVEY.md
§8(f) rows 1 and 3), kept on the host; batches are handed to the CUDA engine as index / feature tensors.
"""
from __future__ import annotations

import os
from dataclasses import dataclass

import numpy as np

_ROOT = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "datasets"))


@dataclass
class Graph:
    label: int
    neighbors: list            # list of int arrays (undirected, deduplicated, no self loops removed)
    node_tags: list
    node_features: np.ndarray = None

    @property
    def n(self):
        return len(self.neighbors)


def load_data(dataset, degree_as_tag=False, root=None):
    """-> (graphs, num_classes).  `dataset` is a name under `root` (default: <repo>/datasets) or a file path."""
    path = dataset if os.path.isfile(dataset) else os.path.join(root or _ROOT, dataset, dataset + ".txt")
    graphs, label_map, tag_map = [], {}, {}
    with open(path) as f:
        n_g = int(f.readline().strip())
        for _ in range(n_g):
            n, l = (int(w) for w in f.readline().split())
            label_map.setdefault(l, len(label_map))
            nbr_sets = [set() for _ in range(n)]
            tags = []
            for j in range(n):
                row = f.readline().split()
                deg = int(row[1])
                tag = int(row[0])
                tag_map.setdefault(tag, len(tag_map))
                tags.append(tag_map[tag])
                for k in row[2:2 + deg]:
                    k = int(k)
                    nbr_sets[j].add(k)
                    nbr_sets[k].add(j)
            graphs.append(Graph(label_map[l], [np.array(sorted(s), dtype=np.int64) for s in nbr_sets], tags))
    if degree_as_tag:
        for g in graphs:
            g.node_tags = [len(nb) for nb in g.neighbors]
    tagset = sorted({t for g in graphs for t in g.node_tags})
    index = {t: i for i, t in enumerate(tagset)}
    for g in graphs:
        feat = np.zeros((g.n, len(tagset)), dtype=np.float32)
        feat[np.arange(g.n), [index[t] for t in g.node_tags]] = 1.0
        g.node_features = feat
    return graphs, len(label_map)


def separate_data(graphs, fold_idx, seed=0):
    """Same folds as the reference (util.py:160-172)."""
    from sklearn.model_selection import StratifiedKFold
    assert 0 <= fold_idx < 10
    skf = StratifiedKFold(n_splits=10, shuffle=True, random_state=seed)
    labels = [g.label for g in graphs]
    train_idx, test_idx = list(skf.split(np.zeros(len(labels)), labels))[fold_idx]
    return [graphs[i] for i in train_idx], [graphs[i] for i in test_idx]


def build_batch(batch_graphs, num_neighbors, rng=np.random, reddit_tile=None):
    """Host batch in the reference's format: input_x[N, k+1] = [node, k neighbours sampled with replacement]
    (isolated nodes repeat themselves), CSR rowptr of the pooling operator, features, labels."""
    sizes = [g.n for g in batch_graphs]
    start = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    X = np.concatenate([g.node_features for g in batch_graphs], 0)
    if reddit_tile:
        X = np.tile(X, reddit_tile) * 0.01
    rows = []
    for gi, g in enumerate(batch_graphs):
        for i, nb in enumerate(g.neighbors):
            node = start[gi] + i
            if len(nb):
                rows.append(np.concatenate([[node], start[gi] + rng.choice(nb, num_neighbors, replace=True)]))
            else:
                rows.append(np.full(num_neighbors + 1, node, dtype=np.int64))
    input_x = np.stack(rows).astype(np.int64)
    labels = np.array([g.label for g in batch_graphs], dtype=np.int64)
    return input_x, start, np.ascontiguousarray(X, dtype=np.float32), labels


def global_node_ids(graphs, selected):
    """input_y of the unsupervised script (train_pytorch_U2GNN_UnSup.py:96-99): dataset-wide node ids."""
    offs = np.concatenate([[0], np.cumsum([g.n for g in graphs])]).astype(np.int64)
    return np.concatenate([np.arange(offs[i], offs[i + 1]) for i in selected]).astype(np.int64)


# ------------------------------------------------------------------------------------------------------------------
# device-side batch builder
# ------------------------------------------------------------------------------------------------------------------
def dataset_csr(graphs):
    """Whole dataset as ONE CSR over dataset-wide node ids -> (rowptr int64 [V+1], col int64 [E], graph_start int64 [G+1],
    features float32 [V, d]).  Neighbour lists keep the order of `Graph.neighbors` (sorted, as the reference's
    edge_mat-derived dict does after its own sort)."""
    sizes = np.array([g.n for g in graphs], dtype=np.int64)
    gstart = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    deg = np.concatenate([[len(nb) for nb in g.neighbors] for g in graphs]).astype(np.int64) if graphs else np.zeros(0, np.int64)
    rowptr = np.concatenate([[0], np.cumsum(deg)]).astype(np.int64)
    parts = [np.asarray(nb, dtype=np.int64) + gstart[gi] for gi, g in enumerate(graphs) for nb in g.neighbors if len(nb)]
    col = np.concatenate(parts).astype(np.int64) if parts else np.zeros(0, np.int64)
    X = np.concatenate([g.node_features for g in graphs], 0).astype(np.float32)
    return rowptr, col, gstart, X


class DeviceBatchBuilder:
    """Keeps the dataset (CSR adjacency + features) in HBM and builds each batch with the CUDA kernels
    `u2gnn_build_batch` (neighbour sampling, index translation) and `u2gnn_gather_rows` (X_concat): what
    `get_batch_data` does on the host in the reference (train_pytorch_U2GNN_Sup.py:91-119).  The only host work per
    batch is the prefix sum over the selected graphs' sizes.  Draws come from the engine's counter-based stream
    (seed, stream = step), not from numpy's generator: use `build_batch` above to reproduce reference index streams."""

    def __init__(self, graphs, num_neighbors, device="cuda", seed=123):
        import torch
        from ._lib import LIB, require_device
        require_device()
        self._torch, self._lib = torch, LIB
        rowptr, col, gstart, X = dataset_csr(graphs)
        self.k, self.seed, self.step = int(num_neighbors), int(seed), 0
        self.gstart_host = gstart
        self.labels_host = np.array([g.label for g in graphs], dtype=np.int64)
        dev = torch.device(device)
        self.rowptr = torch.from_numpy(rowptr).to(dev)
        self.col = torch.from_numpy(col if len(col) else np.zeros(1, np.int64)).to(dev)
        self.X = torch.from_numpy(X).to(dev)
        self.device = dev

    def build(self, selected, stream_id=None):
        """selected: iterable of graph indices -> (input_x [N, k+1] int64, pool rowptr [G+1] int64, X_concat [N, d] f32,
        labels [G] int64, node_global [N] int64), all on the device."""
        torch, LIB = self._torch, self._lib
        from .engine import _ptr, _stream
        sel = np.asarray(list(selected), dtype=np.int64)
        sizes = self.gstart_host[sel + 1] - self.gstart_host[sel]
        off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
        N, G = int(off[-1]), len(sel)
        host = torch.from_numpy(np.concatenate([self.gstart_host[sel], off])).pin_memory()
        meta = host.to(self.device, non_blocking=True)
        gstart_d, off_d = meta[:G], meta[G:]
        input_x = torch.empty((N, self.k + 1), dtype=torch.int64, device=self.device)
        node_global = torch.empty((N,), dtype=torch.int64, device=self.device)
        sid = self.step if stream_id is None else int(stream_id)
        self.step += 1
        LIB.call("u2gnn_build_batch", _ptr(self.rowptr), _ptr(self.col), _ptr(gstart_d), _ptr(off_d), G, N, self.k, self.seed,
                 sid & 0xFFFFFFFF, _ptr(input_x), _ptr(node_global), _stream())
        d = self.X.shape[1]
        Xc = torch.empty((N, d), dtype=torch.float32, device=self.device)
        LIB.call("u2gnn_gather_rows", _ptr(self.X), self.X.shape[0], d, _ptr(node_global), N, 1, _ptr(Xc), _stream())
        labels = torch.from_numpy(self.labels_host[sel]).to(self.device, non_blocking=True)
        return input_x, off_d, Xc, labels, node_global
