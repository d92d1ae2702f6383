"""World-size-2 checks of the data-parallel host logic on CPU (gloo): graph sharding with re-based node
ids, gradient all-reduce, and the row-sharded class-table bookkeeping.  The kernels themselves are covered
by the -m gpu tests; here the per-rank arithmetic is emulated with the numpy oracle."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import load_golden, split_case


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path[:0] = [root, os.path.join(root, "graph-transformer_b200")]
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from u2gnn_b200 import parallel as P
    from oracle import u2gnn_oracle as O
    c = load_golden("sup_neighbors_L2")
    params, grads, _ = split_case(c)
    k, d, ff, T, L, C = [int(v) for v in c["meta"]]
    rowptr = torch.from_numpy(c["rowptr"])
    sh = P.shard_graph_batch(torch.from_numpy(c["input_x"]), rowptr, torch.from_numpy(c["X"]), torch.from_numpy(c["labels"]),
                             rank, world)
    # local forward/backward with the oracle; the loss is a mean over ALL graphs, so local gradients are
    # scaled by G_local / G_total and summed by the all-reduce
    Pm = {n: v.astype(np.float64) for n, v in params.items()}
    G_total = len(c["labels"])
    X = sh["X"].numpy().astype(np.float64)
    scores, cache = O.sup_forward(Pm, sh["input_x"].numpy(), sh["rowptr"].numpy(), X, L, T, "neighbors")
    soft = O.label_smoothing(sh["labels"].numpy(), C, dtype=np.float64)
    loss, dscores = O.soft_cross_entropy(scores, soft)
    G_local = len(sh["labels"])
    dscores = dscores * (G_local / G_total)
    g = O.sup_backward(dscores, cache, Pm, sh["input_x"].numpy(), sh["rowptr"].numpy(), X)
    names = sorted(g)
    flat = torch.from_numpy(np.concatenate([g[n].reshape(-1) for n in names]))
    P.all_reduce_sum_(flat)
    ref = np.concatenate([grads[n].reshape(-1) for n in names])
    err = float(np.abs(flat.numpy() - ref).max() / np.abs(ref).max())
    # row-sharded class table: owner-filled buffers summed by all-reduce reproduce W[ids]
    V, D = 101, 4
    W = torch.arange(V * D, dtype=torch.float32).view(V, D)
    rs = P.RowShard(V, world, rank)
    ids = torch.tensor([0, 3, 50, 51, 100, 77, 3])
    buf = torch.zeros(len(ids), D)
    mine = rs.owns(ids)
    buf[mine] = W[rs.lo:rs.hi][rs.to_local(ids[mine])]
    P.all_reduce_sum_(buf)
    ok_rows = bool(torch.equal(buf, W[ids]))
    q.put((rank, err, ok_rows, sh["node_range"], sh["graph_range"], rs.local_rows))
    dist.destroy_process_group()


def test_world2_sharding_allreduce_and_rowshard():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, e0, ok0, n0, g0, l0), (r1, e1, ok1, n1, g1, l1) = res
    assert e0 < 1e-6 and e1 < 1e-6            # summed shard gradients == gradients of the whole batch (reference fixture)
    assert ok0 and ok1
    assert n0[0] == 0 and n0[1] == n1[0] and g0[1] == g1[0] and g0[0] == 0     # contiguous, disjoint, complete
    assert l0 + l1 == 101


def test_balanced_ranges_cover_everything():
    from u2gnn_b200 import parallel as P
    rowptr = torch.tensor([0, 5, 6, 30, 31, 40, 100, 101])
    for w in (1, 2, 3, 4, 8):
        r = P.balanced_graph_ranges(rowptr, w)
        assert r[0][0] == 0 and r[-1][1] == 7
        assert all(r[i][1] == r[i + 1][0] for i in range(w - 1))
