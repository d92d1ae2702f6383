"""tcgen05 / TMEM path: hardware layout self-tests and the fused bf16 FFN kernels against the fp32
CUDA path and the oracle (tolerance 2e-2, BASELINE.json north_star "bf16-FFN mode")."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def U():
    import u2gnn_b200
    u2gnn_b200.require_device()
    return u2gnn_b200


def bf16_round(x):
    return x.to(torch.bfloat16).to(torch.float32)


@pytest.mark.parametrize("mode,K,N", [(0, 64, 64), (0, 64, 128), (0, 128, 128), (0, 128, 64), (1, 128, 64),
                                      (1, 128, 128), (1, 32, 64), (2, 64, 128), (2, 128, 64), (3, 64, 128), (3, 128, 64)])
def test_tcgen05_operand_paths(U, mode, K, N):
    from u2gnn_b200 import engine as E
    g = torch.Generator(device="cuda").manual_seed(mode * 1000 + K + N)
    if mode == 1:
        A = torch.randn(K, 128, device="cuda", generator=g)      # At[K, M]
        B = torch.randn(K, N, device="cuda", generator=g)        # Bt[K, N]
        ref = bf16_round(A).t() @ bf16_round(B)
    else:
        A = torch.randn(128, K, device="cuda", generator=g)
        B = torch.randn(N, K, device="cuda", generator=g)
        ref = bf16_round(A) @ bf16_round(B).t()
    C = torch.zeros(128, N, device="cuda")
    scratch = torch.zeros(65536, dtype=torch.uint8, device="cuda")
    U.LIB.call("u2gnn_tc_selftest", mode, A.data_ptr(), B.data_ptr(), C.data_ptr(), K, N, scratch.data_ptr(), E._stream())
    torch.cuda.synchronize()
    err = (C - ref).abs().max().item() / ref.abs().max().item()
    assert err < 1e-5, (mode, K, N, err)
