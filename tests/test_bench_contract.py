"""bench.py contract checks that need no GPU: the reference arm never loads the product library, and the torch.nn port it
times (oracle/torch_port.py) reproduces the reference-generated fixtures (same parameters -> same scores / loss / gradients)."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT, load_golden, split_case


def test_reference_arm_does_not_load_the_cuda_package():
    code = ("import sys, io, contextlib, json; sys.argv=['bench.py','--impl','reference','--workload','cfg1','--steps','1','--warmup','1'];"
            "import bench; buf=io.StringIO();\n"
            "with contextlib.redirect_stdout(buf): bench.main()\n"
            "line=json.loads(buf.getvalue().strip().splitlines()[-1]);"
            "maps=open('/proc/self/maps').read();"
            "print(json.dumps({'impl': line['impl'], 'pkg': 'u2gnn_b200' in sys.modules, 'so': 'libu2gnn_b200' in maps, 'value': line['value']}))")
    r = subprocess.run([sys.executable, "-c", code], cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    out = json.loads(r.stdout.strip().splitlines()[-1])
    assert out["impl"] == "reference" and out["value"] > 0
    assert out["pkg"] is False and out["so"] is False


@pytest.mark.parametrize("case", ["sup_neighbors_small", "sup_nodes_L2", "sup_cfg1_shape"])
def test_torch_port_equals_reference_fixture_supervised(case):
    from oracle import torch_port as TP
    c = load_golden(case)
    params, grads, _ = split_case(c)
    k, d, ff, T, L, C = [int(v) for v in c["meta"]]
    m = TP.SupPort(d, ff, C, T, 0.5, L, attn_axis=str(c["attn_axis"]))
    m.load_state_dict({n: torch.from_numpy(v) for n, v in params.items()})
    N, G = c["X"].shape[0], len(c["labels"])
    gp = torch.sparse_coo_tensor(torch.from_numpy(c["pool_idx"]), torch.ones(N), (G, N))
    ix, X = torch.from_numpy(c["input_x"]), torch.from_numpy(c["X"])
    m.eval()
    with torch.no_grad():
        s = m(ix, gp, X).numpy()
    assert np.abs(s - c["eval_scores"]).max() <= 1e-6 * np.abs(c["eval_scores"]).max()
    m.train()
    TP.disable_dropout(m)
    loss = TP.soft_ce(m(ix, gp, X), TP.smooth_labels(torch.from_numpy(c["labels"]), C))
    assert abs(loss.item() - float(c["loss"])) <= 1e-6 * abs(float(c["loss"]))
    loss.backward()
    gmax = max(np.abs(v).max() for v in grads.values())
    for n, p in m.named_parameters():
        assert np.abs(p.grad.numpy() - grads[n]).max() <= 1e-5 * gmax, n


@pytest.mark.parametrize("case", ["unsup_nodes", "unsup_cfg2_shape_nb"])
def test_torch_port_equals_reference_fixture_unsupervised(case):
    from oracle import torch_port as TP
    c = load_golden(case)
    params, grads, _ = split_case(c)
    k, d, ff, T, L, V, ns = [int(v) for v in c["meta"]]
    m = TP.UnSupPort(V, d, ff, T, L, 0.5, attn_axis=str(c["attn_axis"]))
    m.load_state_dict({n: torch.from_numpy(v) for n, v in params.items()})
    m.train()
    TP.disable_dropout(m)
    nl = m(torch.from_numpy(c["X"]), torch.from_numpy(c["input_x"]), torch.from_numpy(c["input_y"]), c["sample_ids"])
    assert np.abs(nl.detach().numpy() - c["node_loss"]).max() <= 1e-5 * np.abs(c["node_loss"]).max()
    torch.sum(nl).backward()
    gmax = max(np.abs(v).max() for v in grads.values())
    for n, p in m.named_parameters():
        assert np.abs(p.grad.numpy() - grads[n]).max() <= 1e-5 * gmax, n
