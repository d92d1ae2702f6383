"""Evaluation path of the two reference scripts, on the CUDA engine.

  supervised    train_pytorch_U2GNN_Sup.py:166-187    test graphs in order, batches of batch_size, neighbours re-sampled from
                                                       the global numpy stream, logits concatenated, argmax accuracy
  unsupervised  train_pytorch_U2GNN_UnSup.py:164-188  graph embeddings = spmm(graph_pool_all, ss.weight) (the CSR segment-sum
                                                       kernel), 10 x StratifiedKFold(seed 0) LogisticRegression(liblinear) fits
The liblinear fits stay on the host, as in the reference; everything that touches the model runs through the C-ABI.
"""
from __future__ import annotations

import numpy as np
import torch

from . import engine as E
from .data import build_batch


class ConditionalStepLR:
    """The reference's scheduler: StepLR(step_size = batches per epoch, gamma 0.1) stepped only after an epoch (> 5) whose
    loss exceeds the mean of the previous five (train_pytorch_U2GNN_Sup.py:147,209-210).  With the default small batches it
    never fires step_size times; with large --batch_size (few steps per epoch) it does, and the lr decays by 10x."""

    def __init__(self, base_lr, step_size, gamma=0.1):
        self.base_lr, self.step_size, self.gamma = float(base_lr), max(int(step_size), 1), float(gamma)
        self.count = 0
        self.losses = []

    def epoch_end(self, loss):
        """-> the learning rate to use from the next epoch on."""
        self.losses.append(float(loss))
        epoch = len(self.losses)
        if epoch > 5 and self.losses[-1] > np.mean(self.losses[-6:-1]):
            self.count += 1
        return self.base_lr * self.gamma ** (self.count // self.step_size)


def sup_logits(model, test_graphs, batch_size, num_neighbors, rng=np.random, reddit_tile=None, device="cuda"):
    """Logits of every test graph, batched exactly like the reference's evaluate()."""
    dev = torch.device(device)
    model.eval()
    out = []
    with torch.no_grad():
        for i in range(0, len(test_graphs), batch_size):
            ix, rp, X, _ = build_batch(test_graphs[i:i + batch_size], num_neighbors, rng, reddit_tile)
            out.append(model(torch.from_numpy(ix).to(dev), torch.from_numpy(rp).to(dev), torch.from_numpy(X).to(dev)))
    return torch.cat(out, 0)


def sup_accuracy(model, test_graphs, batch_size, num_neighbors, rng=np.random, reddit_tile=None, device="cuda"):
    logits = sup_logits(model, test_graphs, batch_size, num_neighbors, rng, reddit_tile, device)
    labels = torch.tensor([g.label for g in test_graphs], device=logits.device)
    return float((logits.argmax(1) == labels).sum().item()) / len(test_graphs)


def unsup_graph_embeddings(weight, pool_rowptr):
    """spmm(graph_pool over ALL graphs, ss.weight): [V, D] -> [G, D] with the segmented-sum kernel (ascending node order)."""
    return E.segment_sum(weight.contiguous(), pool_rowptr)


def unsup_accuracy(weight, pool_rowptr, labels):
    """-> (mean %, std %) of 10-fold logistic regression on the pooled class-table rows."""
    from sklearn.linear_model import LogisticRegression
    from sklearn.model_selection import StratifiedKFold
    emb = unsup_graph_embeddings(weight, pool_rowptr).cpu().numpy()
    labels = np.asarray(labels)
    accs = []
    for fold in range(10):
        skf = StratifiedKFold(n_splits=10, shuffle=True, random_state=0)
        tr, te = list(skf.split(np.zeros(len(labels)), labels))[fold]
        cls = LogisticRegression(solver="liblinear", tol=0.001)
        cls.fit(emb[tr], labels[tr])
        accs.append(cls.score(emb[te], labels[te]))
    return float(np.mean(accs) * 100), float(np.std(accs) * 100)
