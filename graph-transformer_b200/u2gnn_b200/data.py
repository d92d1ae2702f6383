"""Dataset front-end and host batch builder for the CLI / accuracy runs.

Own implementation of what the reference does with networkx in `util.load_data` (U2GNN_pytorch/util.py:54-158)
and `get_batch_data` (train_pytorch_U2GNN_Sup.py:91-119): the text format is parsed straight into per-graph
neighbour lists (numpy), features are the one-hot node tag (or degree, `degree_as_tag`), folds come from the
same `StratifiedKFold(10, shuffle=True, random_state=0)`.  This is plumbing around the hot path (SURVEY.md
§8(f) rows 1 and 3), kept on the host; batches are handed to the CUDA engine as index / feature tensors.

Index parity with the reference (checked bit for bit against reference-generated fixtures in tests/test_data_parity.py):
the reference samples a node's neighbours from `dict_Adj_block`, which lists them in `edge_mat` order
(train_pytorch_U2GNN_Sup.py:100-110).  `edge_mat` is `g.g.edges()` followed by the same pairs reversed (util.py:127-134), and
networkx yields edges node by node in node-INSERTION order, neighbour by neighbour in adjacency-insertion order, skipping
pairs whose other end was already visited.  Node insertion order is the order of first appearance in the file (a node is
created by `add_edge` before its own line is read when an earlier line names it).  `_edge_order` below restates exactly
that; `degree_as_tag` tags follow the same insertion order positionally (util.py:139-141 - a quirk of the reference: row i
gets the degree of the i-th INSERTED node), and the tag -> feature column map is the iteration order of the same Python
set the reference builds (util.py:144-149).
"""
from __future__ import annotations

import os
from dataclasses import dataclass

import numpy as np

_ROOT = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "datasets"))
CACHE_MAGIC = b"U2GNNCSR"
CACHE_VERSION = 1


@dataclass
class Graph:
    label: int
    neighbors: list            # per node: int64 array of neighbour ids in the reference's edge_mat order (a self loop appears twice)
    node_tags: list
    node_features: np.ndarray = None

    @property
    def n(self):
        return len(self.neighbors)


def _edge_order(n, lines):
    """lines[j] = neighbour ids on node j's line.  -> (neighbour lists in edge_mat order, node insertion order, degrees
    in insertion order) with networkx.Graph semantics (util.py:78-100,112-134)."""
    adj = {}                                   # node -> {neighbour: None}, both dicts in insertion order

    def add_node(u):
        if u not in adj:
            adj[u] = {}

    for j in range(n):
        add_node(j)
        for k in lines[j]:
            add_node(k)
            adj[j][k] = None
            adj[k][j] = None
    if len(adj) != n:
        raise ValueError("edge list names a node outside [0, n)")
    seen = set()
    src, dst = [], []
    for u, nbrs in adj.items():                # Graph.edges(): each pair once, from the end inserted first
        for v in nbrs:
            if v not in seen:
                src.append(u)
                dst.append(v)
        seen.add(u)
    out = [[] for _ in range(n)]
    for u, v in zip(src, dst):                 # edge_mat = [edges | reversed edges]; dict_Adj_block[row].append(col) in that order
        out[u].append(v)
    for u, v in zip(src, dst):
        out[v].append(u)
    order = list(adj.keys())
    degree = [len(adj[u]) + (1 if u in adj[u] else 0) for u in order]      # networkx counts a self loop twice
    return [np.array(o, dtype=np.int64) for o in out], order, degree


def _parse_text(path):
    """-> list of (label_raw, neighbour lists, tags_raw, insertion-order degrees)."""
    raw = []
    with open(path) as f:
        n_g = int(f.readline().strip())
        for _ in range(n_g):
            n, l = (int(w) for w in f.readline().split())
            lines, tags = [], []
            for _j in range(n):
                row = f.readline().split()
                deg = int(row[1])
                tags.append(int(row[0]))
                lines.append([int(k) for k in row[2:2 + deg]])
            nbrs, _, degree = _edge_order(n, lines)
            raw.append((l, nbrs, tags, degree))
    return raw


def cache_path(text_path):
    return text_path + ".csr"


def write_cache(path, raw):
    """Binary CSR cache of a parsed dataset: one file, little-endian int64 sections
    [magic 8 B | version | G | V | E] [graph_ptr G+1] [labels G] [tags V] [ins_degree V] [row_ptr V+1] [col E]
    (col = graph-local neighbour ids in the reference's edge_mat order)."""
    G = len(raw)
    sizes = np.array([len(r[1]) for r in raw], dtype=np.int64)
    gptr = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    labels = np.array([r[0] for r in raw], dtype=np.int64)
    tags = np.concatenate([np.asarray(r[2], dtype=np.int64) for r in raw]) if G else np.zeros(0, np.int64)
    insdeg = np.concatenate([np.asarray(r[3], dtype=np.int64) for r in raw]) if G else np.zeros(0, np.int64)
    deg = np.concatenate([[len(nb) for nb in r[1]] for r in raw]).astype(np.int64) if G else np.zeros(0, np.int64)
    rptr = np.concatenate([[0], np.cumsum(deg)]).astype(np.int64)
    parts = [nb for r in raw for nb in r[1] if len(nb)]
    col = np.concatenate(parts).astype(np.int64) if parts else np.zeros(0, np.int64)
    tmp = path + ".tmp.%d" % os.getpid()
    with open(tmp, "wb") as f:
        f.write(CACHE_MAGIC)
        np.array([CACHE_VERSION, G, int(gptr[-1]), len(col)], dtype="<i8").tofile(f)
        for a in (gptr, labels, tags, insdeg, rptr, col):
            a.astype("<i8").tofile(f)
    os.replace(tmp, path)


def read_cache(path):
    with open(path, "rb") as f:
        if f.read(8) != CACHE_MAGIC:
            raise ValueError("%s is not a U2GNN CSR cache" % path)
        ver, G, V, E = (int(v) for v in np.fromfile(f, dtype="<i8", count=4))
        if ver != CACHE_VERSION:
            raise ValueError("%s: cache version %d, expected %d" % (path, ver, CACHE_VERSION))
        gptr = np.fromfile(f, dtype="<i8", count=G + 1)
        labels = np.fromfile(f, dtype="<i8", count=G)
        tags = np.fromfile(f, dtype="<i8", count=V)
        insdeg = np.fromfile(f, dtype="<i8", count=V)
        rptr = np.fromfile(f, dtype="<i8", count=V + 1)
        col = np.fromfile(f, dtype="<i8", count=E)
    if len(col) != E or len(rptr) != V + 1:
        raise ValueError("%s: truncated cache" % path)
    raw = []
    for g in range(G):
        a, b = int(gptr[g]), int(gptr[g + 1])
        nbrs = [col[rptr[v]:rptr[v + 1]].copy() for v in range(a, b)]
        raw.append((int(labels[g]), nbrs, tags[a:b].tolist(), insdeg[a:b].tolist()))
    return raw


def load_data(dataset, degree_as_tag=False, root=None, cache=None):
    """-> (graphs, num_classes).  `dataset` is a name under `root` (default: <repo>/datasets) or a file path.
    cache: None = use `<file>.csr` when it exists and is newer than the text file; True = also (re)write it; False = ignore it."""
    path = dataset if os.path.isfile(dataset) else os.path.join(root or _ROOT, dataset, dataset + ".txt")
    cpath = cache_path(path)
    raw = None
    if cache is not False and os.path.exists(cpath) and os.path.getmtime(cpath) >= os.path.getmtime(path):
        raw = read_cache(cpath)
    if raw is None:
        raw = _parse_text(path)
        if cache:
            write_cache(cpath, raw)
    label_map, tag_map = {}, {}
    graphs = []
    for l, nbrs, tags, degree in raw:
        label_map.setdefault(l, len(label_map))
        mapped = []
        for t in tags:
            tag_map.setdefault(t, len(tag_map))
            mapped.append(tag_map[t])
        graphs.append(Graph(label_map[l], nbrs, degree if degree_as_tag else mapped))
    tagset = set([])
    for g in graphs:                            # same set operations as util.py:144-149 -> same iteration order -> same columns
        tagset = tagset.union(set(g.node_tags))
    tagset = list(tagset)
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
    features float32 [V, d]).  Neighbour lists keep the order of `Graph.neighbors` (the reference's edge_mat order)."""
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

    @classmethod
    def from_device_tensors(cls, rowptr, col, gstart, X, num_neighbors, seed=123, labels=None):
        """Builder over a dataset that is ALREADY resident on the device (synthetic datasets generated in HBM, bench.py cfg4):
        rowptr[V+1], col[E] (dataset-wide CSR), gstart[G+1] (first node of every graph), X[V, d]."""
        import torch
        from ._lib import LIB, require_device
        require_device()
        self = cls.__new__(cls)
        self._torch, self._lib = torch, LIB
        self.k, self.seed, self.step = int(num_neighbors), int(seed), 0
        self.gstart_host = gstart.detach().cpu().numpy().astype(np.int64)
        G = len(self.gstart_host) - 1
        self.labels_host = np.zeros(G, dtype=np.int64) if labels is None else np.asarray(labels, dtype=np.int64)
        self.rowptr, self.col, self.X, self.device = rowptr.contiguous(), col.contiguous(), X.contiguous(), X.device
        return self

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
        LIB.call("u2gnn_gather_rows", _ptr(self.X), self.X.shape[0], d, _ptr(node_global), N, 1, _ptr(Xc), 0, _stream())
        labels = torch.from_numpy(self.labels_host[sel]).to(self.device, non_blocking=True)
        return input_x, off_d, Xc, labels, node_global
