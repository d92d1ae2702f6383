"""Device-side batch builder (reference get_batch_data, train_pytorch_U2GNN_Sup.py:91-119).
CPU: the oracle restatement keeps the reference's invariants.  GPU: the CUDA kernel equals the oracle bit for bit."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "graph-transformer_b200"))
from oracle import u2gnn_oracle as O  # noqa: E402


def _random_dataset(rng, n_graphs, max_nodes, p_edge, isolated_every=5):
    """-> rowptr, col, gstart (dataset-wide CSR, undirected, sorted neighbour lists; some isolated nodes, 1-node graphs)."""
    rows, gstart = [], [0]
    for g in range(n_graphs):
        n = 1 if g % 7 == 3 else int(rng.integers(2, max_nodes + 1))
        A = rng.random((n, n)) < p_edge
        A = np.triu(A, 1)
        A = A | A.T
        if n > 2 and g % isolated_every == 0:
            A[0, :] = False
            A[:, 0] = False
        for i in range(n):
            rows.append(np.nonzero(A[i])[0].astype(np.int64) + gstart[-1])
        gstart.append(gstart[-1] + n)
    deg = np.array([len(r) for r in rows], dtype=np.int64)
    rowptr = np.concatenate([[0], np.cumsum(deg)]).astype(np.int64)
    col = np.concatenate(rows).astype(np.int64) if deg.sum() else np.zeros(0, np.int64)
    return rowptr, col, np.array(gstart, dtype=np.int64)


def _select(rng, gstart, n_sel):
    sel = rng.choice(len(gstart) - 1, n_sel, replace=False).astype(np.int64)
    off = np.concatenate([[0], np.cumsum(gstart[sel + 1] - gstart[sel])]).astype(np.int64)
    return sel, off


@pytest.mark.parametrize("k", [1, 4, 16])
def test_oracle_batch_builder_invariants(k):
    rng = np.random.default_rng(5 + k)
    rowptr, col, gstart = _random_dataset(rng, 40, 30, 0.2)
    sel, off = _select(rng, gstart, 12)
    x, v = O.sample_neighbors_device_stream(rowptr, col, gstart[sel], off, k, 123, 7)
    N = off[-1]
    assert x.shape == (N, k + 1) and x.dtype == np.int64
    assert np.array_equal(x[:, 0], np.arange(N))                       # the node itself first (reference :107)
    for gi, g in enumerate(sel):
        for b in range(off[gi], off[gi + 1]):
            glob = b - off[gi] + gstart[g]
            assert v[b] == glob
            nbrs = col[rowptr[glob]:rowptr[glob + 1]] - gstart[g] + off[gi]
            if len(nbrs) == 0:
                assert np.all(x[b] == b)                               # isolated nodes repeat themselves (:111-112)
            else:
                assert np.isin(x[b, 1:], nbrs).all()                   # only real neighbours, inside the same graph
                assert x[b, 1:].min() >= off[gi] and x[b, 1:].max() < off[gi + 1]
    # a different stream id gives a different draw, the same one the same draw
    x2, _ = O.sample_neighbors_device_stream(rowptr, col, gstart[sel], off, k, 123, 8)
    x3, _ = O.sample_neighbors_device_stream(rowptr, col, gstart[sel], off, k, 123, 7)
    assert np.array_equal(x, x3) and not np.array_equal(x, x2)


def test_oracle_batch_builder_is_uniform():
    # one node with 4 neighbours, many draws: every neighbour ~ 1/4
    rowptr = np.array([0, 4, 5, 6, 7, 8], dtype=np.int64)
    col = np.array([1, 2, 3, 4, 0, 0, 0, 0], dtype=np.int64)
    x, _ = O.sample_neighbors_device_stream(rowptr, col, np.array([0]), np.array([0, 5]), 4096, 9, 1)
    freq = np.bincount(x[0, 1:], minlength=5)[1:] / 4096.0
    assert np.abs(freq - 0.25).max() < 0.03


@pytest.mark.gpu
@pytest.mark.parametrize("n_graphs,max_nodes,k,n_sel", [(40, 30, 4, 12), (300, 140, 16, 200), (9, 5, 8, 9), (64, 600, 16, 64)])
def test_device_batch_builder_matches_oracle(n_graphs, max_nodes, k, n_sel):
    import torch
    import u2gnn_b200 as U
    from u2gnn_b200 import engine as E
    U.require_device()
    rng = np.random.default_rng(n_graphs + k)
    rowptr, col, gstart = _random_dataset(rng, n_graphs, max_nodes, 0.1)
    sel, off = _select(rng, gstart, n_sel)
    N = int(off[-1])
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    x = torch.full((N, k + 1), -1, dtype=torch.int64, device="cuda")
    v = torch.full((N,), -1, dtype=torch.int64, device="cuda")
    colp = dev(col if len(col) else np.zeros(1, np.int64))
    rp, gs, bo = dev(rowptr), dev(gstart[sel]), dev(off)          # keep the device copies alive across the call
    U.LIB.call("u2gnn_build_batch", rp.data_ptr(), colp.data_ptr(), gs.data_ptr(), bo.data_ptr(),
               len(sel), N, k, 0xABCDEF0123456789, 11, x.data_ptr(), v.data_ptr(), E._stream())
    torch.cuda.synchronize()
    xo, vo = O.sample_neighbors_device_stream(rowptr, col, gstart[sel], off, k, 0xABCDEF0123456789, 11)
    assert np.array_equal(x.cpu().numpy(), xo)                         # integer path: bit exact
    assert np.array_equal(v.cpu().numpy(), vo)


@pytest.mark.gpu
def test_device_batch_builder_on_mutag_feeds_the_model():
    import torch
    import u2gnn_b200 as U
    from u2gnn_b200 import data as D
    graphs, C = D.load_data("MUTAG")
    bb = D.DeviceBatchBuilder(graphs, num_neighbors=8, seed=123)
    sel = [3, 50, 7, 120]
    input_x, rowptr, Xc, labels, node_global = bb.build(sel, stream_id=0)
    rp, col, gstart, X = D.dataset_csr(graphs)
    off = np.concatenate([[0], np.cumsum([graphs[i].n for i in sel])]).astype(np.int64)
    xo, vo = O.sample_neighbors_device_stream(rp, col, gstart[np.array(sel)], off, 8, 123, 0)
    assert np.array_equal(input_x.cpu().numpy(), xo) and np.array_equal(node_global.cpu().numpy(), vo)
    assert np.array_equal(rowptr.cpu().numpy(), off)
    assert np.array_equal(Xc.cpu().numpy(), X[vo])                      # gathered features: bit exact
    assert labels.cpu().tolist() == [graphs[i].label for i in sel]
    m = U.TransformerU2GNN(X.shape[1], 64, C, 2, 0.5, 1, attn_axis="neighbors").cuda().eval()
    s = m(input_x, rowptr, Xc)
    assert s.shape == (4, C) and torch.isfinite(s).all()
