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
