#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ by running the REFERENCE itself.

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden.py
Imports /root/reference/U2GNN_pytorch verbatim (pytorch_U2GNN_Sup.TransformerU2GNN,
sampled_softmax.SampledSoftmax, util.load_data, and the train script executed with
--num_epochs 0 for the batch builder).  Two shims, neither touching arithmetic:
  * sys.modules['pyriemann'] stub (imported at util.py:5, used only by get_gm)
  * sys.modules['log_uniform'] backed by the reference C++ sampler class compiled from its own
    sources (oracle/_ref) instead of the Cython binding (log_uniform.pyx).
Writes .npz files; tests/test_oracle_golden.py and the -m gpu tests read them.
"""
import os
import runpy
import sys
import types
import hashlib

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference/U2GNN_pytorch"
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)

from oracle.sampler import RefSampler, build as build_samplers  # noqa: E402

build_samplers(ref=True)
sys.modules["pyriemann"] = types.ModuleType("pyriemann")
_lu = types.ModuleType("log_uniform")
_lu.LogUniformSampler = RefSampler
sys.modules["log_uniform"] = _lu

import pytorch_U2GNN_Sup as ref_sup  # noqa: E402
import sampled_softmax as ref_ss  # noqa: E402


def sd_np(model):
    return {k: v.detach().numpy().copy() for k, v in model.state_dict().items()}


def grads_np(model):
    return {"grad." + k: p.grad.detach().numpy().copy() for k, p in model.named_parameters()}


def no_dropout(model):
    for m in model.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
        if isinstance(m, torch.nn.MultiheadAttention):
            m.dropout = 0.0


def random_batch(rng, sizes, k, d, onehot=True):
    """Synthetic batch in the reference's format: block-diagonal neighbours sampled with
    replacement, isolated nodes repeat themselves (train_pytorch_U2GNN_Sup.py:100-114)."""
    N = int(sum(sizes))
    start = np.concatenate([[0], np.cumsum(sizes)])
    input_x = np.zeros((N, k + 1), dtype=np.int64)
    for g, n in enumerate(sizes):
        for i in range(n):
            node = start[g] + i
            if n == 1 or rng.random() < 0.1:          # isolated node
                input_x[node] = node
            else:
                input_x[node, 0] = node
                input_x[node, 1:] = start[g] + rng.integers(0, n, size=k)
    if onehot:
        X = np.zeros((N, d), dtype=np.float32)
        X[np.arange(N), rng.integers(0, d, size=N)] = 1.0
    else:
        X = rng.standard_normal((N, d)).astype(np.float32)
    idx = np.stack([np.repeat(np.arange(len(sizes)), sizes), np.arange(N)])
    return input_x, X, idx.astype(np.int64), start.astype(np.int64)


def ref_forward(model, input_x, graph_pool, X, attn_axis):
    """nodes: the reference forward verbatim.  neighbors: the same reference modules fed the
    transposed sequence, otherwise the dataflow of pytorch_U2GNN_Sup.py:30-46."""
    if attn_axis == "nodes":
        return model(input_x, graph_pool, X), None
    import torch.nn.functional as F
    scores, outs, src = 0, [], X
    for l in range(model.num_U2GNN_layers):
        seq = F.embedding(input_x, src).transpose(0, 1)
        out = model.u2gnn_layers[l](seq).transpose(0, 1)[:, 0, :]
        outs.append(out)
        ge = model.dropouts[l](torch.spmm(graph_pool, out))
        scores = scores + model.predictions[l](ge)
        src = out
    return scores, outs


def make_sup_case(name, seed, sizes, k, d, ff, T, L, C, attn_axis, onehot=True):
    rng = np.random.default_rng(seed)
    input_x, X, pool_idx, rowptr = random_batch(rng, sizes, k, d, onehot)
    labels = rng.integers(0, C, size=len(sizes)).astype(np.int64)
    torch.manual_seed(seed)
    model = ref_sup.TransformerU2GNN(feature_dim_size=d, ff_hidden_size=ff, num_classes=C,
                                     num_self_att_layers=T, dropout=0.5, num_U2GNN_layers=L)
    # perturb LayerNorm/bias params so gradients wrt them are exercised away from the 1/0 init
    with torch.no_grad():
        for n_, p_ in model.named_parameters():
            if "norm" in n_ or n_.endswith("bias"):
                p_.add_(0.1 * torch.randn_like(p_))
    gp = torch.sparse_coo_tensor(torch.from_numpy(pool_idx), torch.ones(pool_idx.shape[1]),
                                 (len(sizes), int(sum(sizes))))
    tx, tX = torch.from_numpy(input_x), torch.from_numpy(X)
    model.eval()
    with torch.no_grad():
        eval_scores, _ = ref_forward(model, tx, gp, tX, attn_axis)
    model.train()
    no_dropout(model)
    scores, _ = ref_forward(model, tx, gp, tX, attn_axis)
    soft = ref_sup.label_smoothing(torch.from_numpy(labels), C)
    loss = torch.mean(torch.sum(-soft * torch.nn.functional.log_softmax(scores, dim=1), 1))
    loss.backward()
    out = dict(input_x=input_x, X=X, pool_idx=pool_idx, rowptr=rowptr, labels=labels,
               eval_scores=eval_scores.numpy(), train_scores=scores.detach().numpy(),
               soft=soft.numpy(), loss=np.float32(loss.item()),
               meta=np.array([k, d, ff, T, L, C], dtype=np.int64), attn_axis=np.array(attn_axis))
    out.update({"param." + k_: v for k_, v in sd_np(model).items()})
    out.update(grads_np(model))
    # one clip + Adam step (train_pytorch_U2GNN_Sup.py:145,160-161)
    opt = torch.optim.Adam(model.parameters(), lr=5e-4)
    total_norm = torch.nn.utils.clip_grad_norm_(model.parameters(), 0.5)
    opt.step()
    out["grad_norm"] = np.float32(total_norm.item())
    out.update({"after." + k_: v for k_, v in sd_np(model).items()})
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "N", int(sum(sizes)), "loss", loss.item(), "norm", total_norm.item())


def make_unsup_case(name, seed, sizes, k, d, ff, T, L, V, ns, attn_axis):
    """Assembled unsupervised oracle from reference parts: the reference's
    encoder construction + the reference SampledSoftmax class with injected negatives."""
    import torch.nn.functional as F
    from torch.nn import TransformerEncoder, TransformerEncoderLayer
    rng = np.random.default_rng(seed)
    input_x, X, _, _ = random_batch(rng, sizes, k, d, onehot=True)
    N = X.shape[0]
    input_y = rng.choice(V, size=N, replace=False).astype(np.int64)
    torch.manual_seed(seed)
    layers = torch.nn.ModuleList()
    for _ in range(L):  # ctor lines pytorch_U2GNN_UnSup.py:37-40
        enc = TransformerEncoderLayer(d_model=d, nhead=1, dim_feedforward=ff, dropout=0.5)
        layers.append(TransformerEncoder(enc, T))
    ss = ref_ss.SampledSoftmax(V, ns, d * L, torch.device("cpu"))
    ids, true_freq, sample_freq = ss.sampler.sample(ns, input_y)
    no_dropout(layers)
    tx, tX, ty = torch.from_numpy(input_x), torch.from_numpy(X), torch.from_numpy(input_y)
    outs, src = [], tX
    for l in range(L):
        seq = F.embedding(tx, src)
        if attn_axis == "neighbors":
            o = layers[l](seq.transpose(0, 1)).transpose(0, 1)[:, 0, :]
        else:
            o = layers[l](seq)[:, 0, :]
        outs.append(o)
        src = o
    vec = torch.cat(outs, 1)
    logits = ss.sampled(vec, ty, (ids, true_freq, sample_freq))
    loss = torch.sum(logits)
    loss.backward()
    out = dict(input_x=input_x, X=X, input_y=input_y, sample_ids=np.array(ids, dtype=np.int64),
               node_loss=logits.detach().numpy(), loss=np.float32(loss.item()), vec=vec.detach().numpy(),
               meta=np.array([k, d, ff, T, L, V, ns], dtype=np.int64), attn_axis=np.array(attn_axis))
    for k_, v in layers.state_dict().items():
        out["param.u2gnn_layers." + k_] = v.numpy().copy()
    for k_, p in layers.named_parameters():
        out["grad.u2gnn_layers." + k_] = p.grad.numpy().copy()
    out["param.ss.weight"] = ss.weight.detach().numpy().copy()
    out["grad.ss.weight"] = ss.weight.grad.numpy().copy()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "N", N, "loss", loss.item())


def make_mutag_kat():
    """First batch of the reference script on MUTAG (cfg1) + eval-mode scores at init."""
    argv = sys.argv
    sys.argv = ["train_pytorch_U2GNN_Sup.py", "--dataset", "MUTAG", "--fold_idx", "1", "--num_neighbors", "8",
                "--num_timesteps", "3", "--ff_hidden_size", "1024", "--batch_size", "4", "--num_epochs", "0",
                "--run_folder", "/tmp/u2gnn_golden/x/", "--model_name", "MUTAG_kat"]
    cwd = os.getcwd()
    os.chdir(REF)
    try:
        g = runpy.run_path(os.path.join(REF, "train_pytorch_U2GNN_Sup.py"), run_name="__main__")
    finally:
        os.chdir(cwd)
        sys.argv = argv
    input_x, graph_pool, X_concat, graph_labels = g["batch_nodes"]()
    model = g["model"]
    model.eval()
    with torch.no_grad():
        scores = model(input_x, graph_pool, X_concat)
    gp = graph_pool.coalesce() if not graph_pool.is_coalesced() else graph_pool
    out = dict(input_x=input_x.numpy(), X=X_concat.numpy(), pool_idx=graph_pool._indices().numpy(),
               labels=graph_labels.numpy(), eval_scores=scores.numpy(),
               sha1=np.array(hashlib.sha1(input_x.numpy().tobytes()).hexdigest()[:16]),
               n_params=np.int64(sum(p.numel() for p in model.parameters())))
    out.update({"param." + k_: v for k_, v in sd_np(model).items()})
    np.savez_compressed(os.path.join(HERE, "mutag_kat.npz"), **out)
    print("mutag_kat N", input_x.shape[0], "sum", int(input_x.sum()), out["sha1"], scores.numpy().tolist())


def make_data_fixtures():
    """Dataset front-end + host batch builder KATs straight from the reference (util.load_data, get_batch_data): a PTC
    degree-as-tag batch and whole-dataset digests of the edge_mat-order neighbour lists / features / folds."""
    argv = sys.argv
    sys.argv = ["train_pytorch_U2GNN_Sup.py", "--dataset", "MUTAG", "--fold_idx", "1", "--num_neighbors", "4",
                "--num_timesteps", "1", "--ff_hidden_size", "64", "--batch_size", "4", "--num_epochs", "0",
                "--run_folder", "/tmp/u2gnn_golden/x/", "--model_name", "MUTAG_kat2"]
    cwd = os.getcwd()
    os.chdir(REF)
    try:
        g = runpy.run_path(os.path.join(REF, "train_pytorch_U2GNN_Sup.py"), run_name="__main__")
    finally:
        os.chdir(cwd)
        sys.argv = argv
    out = {}
    for name, deg in (("MUTAG", False), ("PTC", True), ("PTC", False)):
        graphs, C = g["load_data"](name, deg)
        key = "%s_%d" % (name, int(deg))
        nbr = []
        for gr in graphs:
            d = {}
            for r, c in zip(gr.edge_mat[0], gr.edge_mat[1]):        # the order dict_Adj_block sees (train_pytorch_U2GNN_Sup.py:100-105)
                d.setdefault(int(r), []).append(int(c))
            for v in range(len(gr.g)):
                nbr.append(np.array(d.get(v, []), dtype=np.int64))
        deg_arr = np.array([len(a) for a in nbr], dtype=np.int64)
        col = np.concatenate(nbr) if len(nbr) else np.zeros(0, np.int64)
        feat = np.concatenate([gr.node_features for gr in graphs], 0)
        out[key + "_deg_sha1"] = np.array(hashlib.sha1(deg_arr.tobytes()).hexdigest())
        out[key + "_col_sha1"] = np.array(hashlib.sha1(col.tobytes()).hexdigest())
        out[key + "_feat_sha1"] = np.array(hashlib.sha1(np.ascontiguousarray(feat, dtype=np.float32).tobytes()).hexdigest())
        out[key + "_labels"] = np.array([gr.label for gr in graphs], dtype=np.int64)
        out[key + "_shape"] = np.array([len(graphs), feat.shape[0], feat.shape[1], len(col), C], dtype=np.int64)
        tr, te = g["separate_data"](graphs, 3)
        ids = {id(gr): i for i, gr in enumerate(graphs)}
        out[key + "_fold3_test"] = np.array([ids[id(gr)] for gr in te], dtype=np.int64)
        if name == "PTC" and deg:
            sel = [5, 100, 7, 343]
            np.random.seed(321)
            ix, gp, X, y = g["get_batch_data"]([graphs[i] for i in sel])     # num_neighbors = 4 from the argv above
            out["ptc_batch_sel"] = np.array(sel, dtype=np.int64)
            out["ptc_batch_input_x"] = ix.numpy()
            out["ptc_batch_X"] = X.numpy()
            out["ptc_batch_labels"] = y.numpy()
            out["ptc_batch_pool_idx"] = gp._indices().numpy()
    np.savez_compressed(os.path.join(HERE, "data_kat.npz"), **out)
    print("data_kat", {k: (v.tolist() if v.ndim and v.size < 8 else str(v)[:18]) for k, v in out.items() if "shape" in k or "sha1" in k})


def make_eval_fixtures():
    """Evaluation path KATs: the reference's supervised evaluate() on the MUTAG fold-1 test graphs at
    init (train_pytorch_U2GNN_Sup.py:166-187) and the unsupervised protocol (spmm over ALL graphs of the class table ->
    10-fold LogisticRegression, train_pytorch_U2GNN_UnSup.py:164-188) on PTC degree-as-tag with a fixed table."""
    from scipy.sparse import coo_matrix  # noqa: F401  (sklearn dependency check)
    from sklearn.linear_model import LogisticRegression
    from sklearn.model_selection import StratifiedKFold
    argv = sys.argv
    sys.argv = ["train_pytorch_U2GNN_Sup.py", "--dataset", "MUTAG", "--fold_idx", "1", "--num_neighbors", "8",
                "--num_timesteps", "3", "--ff_hidden_size", "1024", "--batch_size", "4", "--num_epochs", "0",
                "--run_folder", "/tmp/u2gnn_golden/x/", "--model_name", "MUTAG_eval"]
    cwd = os.getcwd()
    os.chdir(REF)
    try:
        g = runpy.run_path(os.path.join(REF, "train_pytorch_U2GNN_Sup.py"), run_name="__main__")
    finally:
        os.chdir(cwd)
        sys.argv = argv
    model, test_graphs = g["model"], g["test_graphs"]
    with torch.no_grad():                       # move the classifier away from its init so the argmax is not degenerate
        for n_, p_ in model.named_parameters():
            if n_.startswith("predictions"):
                p_.add_(0.05 * torch.randn_like(p_))
    np.random.seed(5)
    acc = g["evaluate"]()
    model.eval()
    np.random.seed(5)
    outs = []
    with torch.no_grad():
        for i in range(0, len(test_graphs), 4):
            ix, gp, X, _ = g["get_batch_data"](test_graphs[i:i + 4])
            outs.append(model(ix, gp, X))
    out = dict(sup_logits=torch.cat(outs, 0).numpy(), sup_acc=np.float64(acc),
               sup_labels=np.array([gr.label for gr in test_graphs], dtype=np.int64))
    out.update({"param." + k_: v for k_, v in sd_np(model).items()})
    # unsupervised: the reference's lines on a fixed, exactly representable class table
    graphs, _ = g["load_data"]("PTC", True)
    V = sum(len(gr.g) for gr in graphs)
    rng = np.random.default_rng(77)
    Wq = rng.integers(-64, 65, size=(V, 4)).astype(np.int8)
    W = torch.from_numpy(Wq.astype(np.float32) / 64.0)
    graph_pool = g["get_graphpool"](graphs)
    labels = np.array([gr.label for gr in graphs])
    emb = torch.spmm(graph_pool, W).data.cpu().numpy()
    accs = []
    for fold in range(10):
        skf = StratifiedKFold(n_splits=10, shuffle=True, random_state=0)
        tr, te = list(skf.split(np.zeros(len(labels)), labels))[fold]
        cls = LogisticRegression(solver="liblinear", tol=0.001)
        cls.fit(emb[tr], labels[tr])
        accs.append(cls.score(emb[te], labels[te]))
    out.update(unsup_Wq=Wq, unsup_emb=emb, unsup_mean=np.float64(np.mean(accs) * 100), unsup_std=np.float64(np.std(accs) * 100))
    np.savez_compressed(os.path.join(HERE, "eval_kat.npz"), **out)
    print("eval_kat sup acc", acc, "unsup", np.mean(accs) * 100, np.std(accs) * 100)


def make_sampler_sets():
    out = {}
    for V, ns in [(100, 50), (3371, 512), (8792, 512), (2540000, 512)]:
        s = RefSampler(V)
        for call in range(2):
            ids, tries = s.sample_with_tries(ns)
            out[f"ids_{V}_{ns}_{call}"] = np.sort(ids)
            out[f"tries_{V}_{ns}_{call}"] = np.int64(tries)
        out[f"prob_{V}"] = np.array([s.probability(i) for i in (0, 1, 7, V - 1)], dtype=np.float32)
        out[f"expcnt_{V}_{ns}"] = s.expected_count(tries, ids[:16])
        out[f"expcnt_ids_{V}_{ns}"] = ids[:16]
    np.savez_compressed(os.path.join(HERE, "sampler_sets.npz"), **out)
    print("sampler_sets", {k: int(v) for k, v in out.items() if k.startswith("tries")})


if __name__ == "__main__":
    if "--data-only" in sys.argv:
        make_data_fixtures()
        sys.exit(0)
    if "--eval-only" in sys.argv:
        make_eval_fixtures()
        sys.exit(0)
    if "--cfg-shapes-only" in sys.argv:
        make_sup_case("sup_cfg1_shape",      31, [17, 23, 13, 28],           8, 7,  1024, 3, 1, 2, "nodes")
        make_unsup_case("unsup_cfg2_shape",  32, [26, 19, 31, 14],           4, 4,  1024, 2, 1, 8792, 512, "nodes")
        make_unsup_case("unsup_cfg2_shape_nb", 33, [26, 19, 31, 14],         4, 4,  1024, 2, 1, 8792, 512, "neighbors")
        sys.exit(0)
    make_mutag_kat()
    make_data_fixtures()
    make_eval_fixtures()
    make_sampler_sets()
    #                 name                    seed sizes                 k  d   ff   T  L  C  axis
    make_sup_case("sup_neighbors_small", 11, [5, 1, 9, 3, 2],            4, 7,  32,  2, 1, 2, "neighbors")
    make_sup_case("sup_neighbors_L2",    12, [6, 4, 11, 2],              8, 12, 48,  3, 2, 3, "neighbors", onehot=False)
    make_sup_case("sup_neighbors_d65",   13, [12, 17, 8],               16, 65, 128, 2, 1, 2, "neighbors")
    make_sup_case("sup_neighbors_d64",   16, [9, 14, 20, 3],            16, 64, 256, 2, 1, 2, "neighbors", onehot=False)
    make_sup_case("sup_nodes_small",     14, [5, 1, 9, 3, 2],            4, 7,  32,  2, 1, 2, "nodes")
    make_sup_case("sup_nodes_L2",        15, [17, 23, 9, 28],            8, 7,  64,  3, 2, 2, "nodes")
    make_unsup_case("unsup_neighbors",   21, [7, 3, 12, 5],              4, 4,  32,  2, 2, 300, 40, "neighbors")
    make_unsup_case("unsup_nodes",       22, [7, 3, 12, 5],              4, 4,  32,  2, 1, 300, 40, "nodes")
    # the exact shapes of BASELINE.json configs[0] / configs[1] (MUTAG: d 7, k 8, T 3, ff 1024; PTC degree-as-tag: d 4, k 4, T 2,
    # ff 1024, V 8 792, ns 512), attention as the reference runs it (nodes) and the intended layout
    make_sup_case("sup_cfg1_shape",      31, [17, 23, 13, 28],           8, 7,  1024, 3, 1, 2, "nodes")
    make_unsup_case("unsup_cfg2_shape",  32, [26, 19, 31, 14],           4, 4,  1024, 2, 1, 8792, 512, "nodes")
    make_unsup_case("unsup_cfg2_shape_nb", 33, [26, 19, 31, 14],         4, 4,  1024, 2, 1, 8792, 512, "neighbors")
