"""CPU port of the reference model on the same stock ``torch.nn`` modules (TEST/BENCH BASELINE ONLY).

The reference's arithmetic for this path lives in PyTorch itself (third-party, not under
/root/reference): ``nn.TransformerEncoder(nn.TransformerEncoderLayer(d, 1, ff, 0.5), T)``,
``F.embedding``, ``torch.spmm`` (pytorch_U2GNN_Sup.py:18-28,30-46).  This port calls exactly
those modules so that (a) the numpy oracle can be checked against autograd anywhere torch is
installed, and (b) ``bench.py`` has the reference's CPU path to time on the GPU box, where
/root/reference does not exist.  It is checked against the reference-generated fixtures (same
parameters -> same scores, loss and gradients) by ``tests/test_bench_contract.py``.

Only tests/, bench.py's cpu_baseline / --impl reference legs and __graft_entry__.smoke() may
import this file.
"""
from __future__ import annotations

import math
import torch
import torch.nn as nn
import torch.nn.functional as F


def build_encoder_stack(d, ff, T, L):
    """Same construction order as pytorch_U2GNN_Sup.py:18-21 -> identical init under a seed."""
    stack = nn.ModuleList()
    for _ in range(L):
        layer = nn.TransformerEncoderLayer(d_model=d, nhead=1, dim_feedforward=ff, dropout=0.5)
        stack.append(nn.TransformerEncoder(layer, T))
    return stack


def run_stack_layer(encoder, src, input_x, attn_axis):
    """gather -> encoder -> sequence position 0 for one U2GNN layer."""
    seq = F.embedding(input_x, src)                      # [N, S, d]
    if attn_axis == "nodes":                             # as written: torch reads (S=N, B=k+1, E=d)
        out = encoder(seq)
    elif attn_axis == "neighbors":                       # intended: (S=k+1, B=N, E=d)
        out = encoder(seq.transpose(0, 1)).transpose(0, 1)
    else:
        raise ValueError(attn_axis)
    return out[:, 0, :]


class SupPort(nn.Module):
    """Supervised TransformerU2GNN (pytorch_U2GNN_Sup.py:7-46) with a selectable attention axis."""

    def __init__(self, feature_dim_size, ff_hidden_size, num_classes, num_self_att_layers, dropout,
                 num_U2GNN_layers, attn_axis="nodes"):
        super().__init__()
        self.attn_axis = attn_axis
        self.L = num_U2GNN_layers
        self.u2gnn_layers = build_encoder_stack(feature_dim_size, ff_hidden_size, num_self_att_layers,
                                                num_U2GNN_layers)
        self.predictions = nn.ModuleList()
        self.dropouts = nn.ModuleList()
        for _ in range(self.L):
            self.predictions.append(nn.Linear(feature_dim_size, num_classes))
            self.dropouts.append(nn.Dropout(dropout))

    def forward(self, input_x, graph_pool, X_concat):
        scores = 0
        src = X_concat
        for l in range(self.L):
            out = run_stack_layer(self.u2gnn_layers[l], src, input_x, self.attn_axis)
            ge = self.dropouts[l](torch.spmm(graph_pool, out))
            scores = scores + self.predictions[l](ge)
            src = out
        return scores


def soft_ce(pred, soft):
    """train_pytorch_U2GNN_Sup.py:140-142."""
    return torch.mean(torch.sum(-soft * F.log_softmax(pred, dim=1), 1))


def smooth_labels(labels, classes, smoothing=0.1):
    """pytorch_U2GNN_Sup.py:48-59."""
    t = torch.full((labels.numel(), classes), smoothing / (classes - 1))
    t.scatter_(1, labels.view(-1, 1), 1.0 - smoothing)
    return t


class SampledSoftmaxPort(nn.Module):
    """sampled_softmax.py:11-56 with the negatives injected (the ``sampled`` entry point)."""

    def __init__(self, ntokens, nhid):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(ntokens, nhid))
        stdv = math.sqrt(6.0 / (ntokens + nhid))
        self.weight.data.uniform_(-stdv, stdv)

    def sampled(self, inputs, labels, sample_ids):
        ids = torch.as_tensor(sample_ids, dtype=torch.long)
        tw = self.weight.index_select(0, labels)
        sw = self.weight.index_select(0, ids)
        true_e = torch.exp((inputs * tw).sum(1))
        samp_e = torch.exp(inputs @ sw.t())
        return -torch.log(true_e / samp_e.sum(1))


class UnSupPort(nn.Module):
    """Assembled unsupervised model (ctor lines pytorch_U2GNN_UnSup.py:37-44)."""

    def __init__(self, vocab_size, feature_dim_size, ff_hidden_size, num_self_att_layers,
                 num_U2GNN_layers, dropout, attn_axis="nodes"):
        super().__init__()
        self.attn_axis = attn_axis
        self.L = num_U2GNN_layers
        self.u2gnn_layers = build_encoder_stack(feature_dim_size, ff_hidden_size, num_self_att_layers,
                                                num_U2GNN_layers)
        self.dropouts = nn.Dropout(dropout)
        self.ss = SampledSoftmaxPort(vocab_size, feature_dim_size * num_U2GNN_layers)

    def forward(self, X_concat, input_x, input_y, sample_ids):
        outs = []
        src = X_concat
        for l in range(self.L):
            out = run_stack_layer(self.u2gnn_layers[l], src, input_x, self.attn_axis)
            outs.append(out)
            src = out
        vec = self.dropouts(torch.cat(outs, 1))
        return self.ss.sampled(vec, input_y, sample_ids)


def disable_dropout(model):
    """p=0 everywhere (incl. the attention-prob dropout inside MHA) for gradient parity."""
    for m in model.modules():
        if isinstance(m, nn.Dropout):
            m.p = 0.0
        if isinstance(m, nn.MultiheadAttention):
            m.dropout = 0.0
    return model


def train_step(model, opt, loss_fn):
    """forward + loss + backward + clip 0.5 + Adam (train_pytorch_U2GNN_Sup.py:155-161)."""
    opt.zero_grad()
    loss = loss_fn()
    loss.backward()
    torch.nn.utils.clip_grad_norm_(model.parameters(), 0.5)
    opt.step()
    return loss
