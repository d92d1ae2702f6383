"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/u2gnn_b200.h declares, validates arguments without a GPU, and the host-side logic
(dropout stream ids, flat parameter arena, module surface) is consistent with the oracle."""
import ctypes
import os
import subprocess

import numpy as np
import pytest
import torch

from conftest import ROOT
from oracle import u2gnn_oracle as O


@pytest.fixture(scope="module")
def U():
    import u2gnn_b200
    return u2gnn_b200


def test_library_exports_every_declared_symbol(U):
    assert os.path.exists(U.LIB_PATH)
    syms = subprocess.run(["nm", "-D", "--defined-only", U.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = {l.split()[-1] for l in syms.splitlines() if " T " in l}
    declared = set(U.SIGNATURES)
    assert len(declared) >= 30
    assert declared <= exported, declared - exported
    assert {e for e in exported if e.startswith("u2gnn_")} == declared     # nothing undeclared leaks out


def test_product_library_has_no_probe_entry_points_or_global_state(U):
    """probes / self-tests / tracing live in libu2gnn_b200_probe.so; the product library exports only what
    include/u2gnn_b200.h declares and defines no writable global data of its own (B / D symbols)."""
    from u2gnn_b200 import _lib
    syms = subprocess.run(["nm", "-D", "--defined-only", U.LIB_PATH], capture_output=True, text=True, check=True).stdout
    names = [l.split()[-1] for l in syms.splitlines()]
    assert not [n for n in names if any(w in n for w in ("probe", "selftest", "trace", "debug", "_mode"))]
    assert not [l for l in syms.splitlines() if l.split()[-2] in ("B", "D") and l.split()[-1].startswith(("g_", "u2gnn"))]
    probe_decl = set(_lib.parse_header(_lib.PROBE_HEADER_PATH))
    assert probe_decl and not (probe_decl & set(U.SIGNATURES))
    psyms = subprocess.run(["nm", "-D", "--defined-only", _lib.PROBE_LIB_PATH], capture_output=True, text=True, check=True).stdout
    pexp = {l.split()[-1] for l in psyms.splitlines() if " T " in l}
    assert probe_decl <= pexp and set(U.SIGNATURES) <= pexp       # the probe library is a superset build


def test_header_has_no_torch_types_and_cites_reference():
    text = open(os.path.join(ROOT, "include", "u2gnn_b200.h")).read()
    assert "at::" not in text and "torch::" not in text and "#include <torch" not in text and "Tensor " not in text
    for cite in ("pytorch_U2GNN_Sup.py:32", "sampled_softmax.py:36-56", "Log_Uniform_Sampler.cpp:57-71",
                 "train_pytorch_U2GNN_Sup.py:145"):
        assert cite in text


def test_status_codes_without_gpu(U):
    lib = U.LIB
    assert lib.cdll.u2gnn_version() >= 100
    assert "invalid" in lib.strerror(-1) and "workspace" in lib.strerror(-5)
    if not torch.cuda.is_available():
        assert lib.cdll.u2gnn_device_check() == -6
        with pytest.raises(RuntimeError):
            U.require_device()
    # argument validation happens before any launch: null pointers / bad sizes -> EINVAL, raised as RuntimeError
    with pytest.raises(RuntimeError, match="invalid argument"):
        lib.call("u2gnn_gather_rows", 0, 10, 4, 0, 5, 1, 0, 0, 0)
    with pytest.raises(RuntimeError, match="invalid argument"):
        lib.call("u2gnn_seqattn_fwd", 1, 4, 40, 40, 8, 0, 0, 0, 1, 0)      # S > 32
    with pytest.raises(RuntimeError, match="invalid argument"):
        lib.call("u2gnn_logu_sample", 10, 11, 1, 1, 1, 1, 1 << 20, 0)      # size > N: reference would spin forever
    assert lib.call("u2gnn_logu_sample_workspace_bytes", 512) >= 8 * (512 + 2048)


def test_host_rng_word_matches_oracle(U):
    for p in (0.5, 0.25, 0.1):
        thr = O.dropout_threshold(p)
        keep, _ = O.dropout_keep_mask(0x1234567890ABCDEF, 21, 64, p)
        for g in range(2):
            w = U.LIB.call("u2gnn_rng_mask_word_host", 0x1234567890ABCDEF, 21, g, thr)
            assert [(w >> i) & 1 for i in range(32)] == [int(b) for b in keep[32 * g:32 * g + 32]]


def test_stream_ids_agree_with_oracle(U):
    from u2gnn_b200 import engine as E
    for l in range(3):
        for t in range(4):
            for s in range(4):
                assert E.stream_id(l, t, s, 4) == O.stream_id(l, t, s, 4)
    assert (E.STREAM_POOLED, E.STREAM_CONCAT) == (O.STREAM_POOLED, O.STREAM_CONCAT)
    assert E.dropout_threshold(0.5) == 128 == O.dropout_threshold(0.5)


def test_module_surface_and_state_dict_names(U):
    torch.manual_seed(123)
    m = U.TransformerU2GNN(feature_dim_size=7, ff_hidden_size=1024, num_classes=2, dropout=0.5,
                           num_self_att_layers=3, num_U2GNN_layers=1)
    assert sum(p.numel() for p in m.parameters()) == 46873
    keys = set(m.state_dict())
    assert "u2gnn_layers.0.layers.2.self_attn.in_proj_weight" in keys and "predictions.0.bias" in keys
    from oracle.torch_port import SupPort
    torch.manual_seed(123)
    ref = SupPort(7, 1024, 2, 3, 0.5, 1)
    for (n1, p1), (n2, p2) in zip(m.state_dict().items(), ref.state_dict().items()):
        assert n1 == n2 and torch.equal(p1, p2)
    u = U.TransformerU2GNNUnSup(vocab_size=50, feature_dim_size=4, ff_hidden_size=16, sampled_num=8,
                                num_self_att_layers=2, num_U2GNN_layers=2, dropout=0.5, device=torch.device("cpu"))
    assert u.ss.weight.shape == (50, 8)
    with pytest.raises(RuntimeError):                                       # product path fails loudly without a GPU
        if torch.cuda.is_available():
            raise RuntimeError("gpu present")
        m(torch.zeros(3, 5, dtype=torch.int64), torch.zeros(2, dtype=torch.int64), torch.zeros(3, 7))


def test_flat_arena_views_keep_state_dict(U):
    from u2gnn_b200.trainer import FlatArena
    torch.manual_seed(0)
    m = U.TransformerU2GNN(5, 8, 2, 1, 0.5, 2)
    before = {k: v.clone() for k, v in m.state_dict().items()}
    a = FlatArena(m)
    assert a.total % 4 == 0 and a.total >= sum(p.numel() for p in m.parameters())
    for k, v in m.state_dict().items():
        assert torch.equal(v, before[k])
    a.p.mul_(2.0)
    assert torch.equal(m.predictions[0].weight.data, before["predictions.0.weight"] * 2.0)


def test_tied_timesteps_share_one_weight_set_and_one_gradient_slot(U):
    """tie_timesteps=True: T timesteps alias ONE encoder layer; state_dict keeps the reference names,
    the flat arena holds one copy and every timestep's gradient dict points at the same slot."""
    from u2gnn_b200.trainer import FlatArena
    from u2gnn_b200.model import _layer_param_dicts
    torch.manual_seed(1)
    free = U.TransformerU2GNN(6, 16, 2, 3, 0.5, 1)
    tied = U.TransformerU2GNN(6, 16, 2, 3, 0.5, 1, tie_timesteps=True)
    assert set(tied.state_dict()) == set(free.state_dict())
    per_layer = sum(p.numel() for p in free.u2gnn_layers[0].layers[0].parameters())
    assert sum(p.numel() for p in free.parameters()) - sum(p.numel() for p in tied.parameters()) == 2 * per_layer
    sd = free.state_dict()
    for t in (1, 2):                                     # a reference checkpoint with identical copies loads unchanged
        for k in list(sd):
            if ".layers.%d." % t in k:
                sd[k] = sd[k.replace(".layers.%d." % t, ".layers.0.")]
    tied.load_state_dict(sd)
    a = FlatArena(tied)
    g = a.grad_dicts([_layer_param_dicts(tied.u2gnn_layers[0])])
    assert len(g[0]) == 3
    for n in g[0][0]:
        assert g[0][0][n].data_ptr() == g[0][1][n].data_ptr() == g[0][2][n].data_ptr()
    u = U.TransformerU2GNNUnSup(vocab_size=50, feature_dim_size=4, ff_hidden_size=16, sampled_num=8, num_self_att_layers=2,
                                num_U2GNN_layers=1, dropout=0.5, device=torch.device("cpu"), tie_timesteps=True)
    assert u.u2gnn_layers[0].layers[0] is u.u2gnn_layers[0].layers[1]


def test_label_smoothing_matches_oracle(U):
    y = torch.tensor([0, 2, 1, 1])
    assert np.allclose(U.label_smoothing(y, 3).numpy(), O.label_smoothing(y.numpy(), 3))


def test_tf_style_label_smoothing_is_the_reference_formula_with_rescaled_eps(U):
    """tf.losses.softmax_cross_entropy(label_smoothing=eps): onehot * (1 - eps) + eps / C  ==  the reference's targets with
    eps' = eps (C - 1) / C (trainer.effective_smoothing), so the loss kernel needs no second path."""
    from u2gnn_b200.trainer import effective_smoothing
    labels = torch.tensor([0, 2, 1, 2])
    for C in (2, 3, 7):
        lab = labels % C
        eps = 0.1
        tf_targets = torch.nn.functional.one_hot(lab, C).float() * (1 - eps) + eps / C
        ours = U.label_smoothing(lab, C, smoothing=effective_smoothing(eps, C, "tf"))
        assert torch.allclose(ours, tf_targets, atol=1e-7)
        assert effective_smoothing(eps, C, "reference") == eps
    with pytest.raises(ValueError):
        effective_smoothing(0.1, 2, "other")


def test_ffn_epilogue_bit_arithmetic_on_the_host():
    """csrc/ffn_epi.cuh, forward chunk epilogue: dropout is a bf16 multiplication there - the keep bits of register pair j (elements
    2j, 2j + 1 of a 32-element group) must land at bits 14 / 30 of a word (the bf16 pair 2.0 | 0.0 per half) through ONE integer
    multiplication of a pre-masked word and one AND, with no carry between the two shifted copies; and the mask words the three FFN
    kernels exchange use the order in which the sign bytes of two g registers land (flag_pos / flag_elem are inverse permutations with
    both halves of a pair 16 bits apart).  The same inline functions, compiled for the host in the probe library."""
    from u2gnn_b200._lib import probe_lib
    P = probe_lib()
    kp = (ctypes.c_uint32 * 16)()
    pos = (ctypes.c_int * 32)()
    elem = (ctypes.c_int * 32)()
    rng = np.random.RandomState(5)
    words = [0, 0xFFFFFFFF, 0x55555555, 0xAAAAAAAA, 0x0000FFFF, 0xFFFF0000, 0x00008000, 0x00004000, 0x80000000, 0xC000C000, 0x3FFF3FFF]
    words += [int(w) for w in rng.randint(0, 2 ** 32, size=2000, dtype=np.uint64)]
    for w in words:
        assert P.call("u2gnn_epi_bits_host", w, ctypes.addressof(kp), ctypes.addressof(pos), ctypes.addressof(elem)) == 0
        for j in range(16):
            want = (0x4000 if (w >> (2 * j)) & 1 else 0) | (0x40000000 if (w >> (2 * j + 1)) & 1 else 0)
            assert kp[j] == want, (hex(w), j, hex(kp[j]), hex(want))
    p, e = list(pos), list(elem)
    assert sorted(p) == list(range(32)) and all(e[p[i]] == i for i in range(32))
    for j in range(16):
        assert p[2 * j + 1] == p[2 * j] + 16 and p[2 * j] == 8 * (j & 1) + (j >> 1)
