// Fused bf16 FFN block on tcgen05 + TMEM, backward entry point.  The [rows, ff] hidden activation and its
// gradient are recomputed on chip (never stored in HBM), so the backward is two kernels with opposite
// loop orders (DESIGN.md "FFN backward"):
//
//   dgrad  (ffn_tc_dgrad.cu: rows outer, ff chunks inner)          dy1 = dz + dPre W1
//   wgrad  (ffn_tc_wgrad.cu: one ff chunk per CTA, row tiles inner)  dW1, db1, dW2
//
// with X = y1 (block input), dF = gradient at the linear2 output (after the output dropout),
// dz = gradient at the residual sum.  These are the autograd of linear1/ReLU/dropout/linear2 in
// nn.TransformerEncoderLayer._ff_block (torch/nn/modules/transformer.py:977-982).
#include "common.cuh"

int ffn_tc_dgrad_launch(const float* y1, const float* df, const float* dz, float* dy1, int64_t M, int d, int ff,
                        const void* packed, uint64_t seed, uint32_t stream_hidden, int thr, void* xb, void* fb, cudaStream_t st);
int ffn_tc_wgrad_launch(const void* xb, const void* fb, int64_t M, int d, int ff, const void* packed, float hidden_scale,
                        uint64_t seed, uint32_t stream_hidden, int thr, float* dW1, float* db1, float* dW2, cudaStream_t st);

extern "C" size_t u2gnn_ffn_tc_bwd_workspace_bytes(int64_t M) {
    if (M < 0) return 0;
    return (size_t)(2 * ((M + 255) / 256)) * 2 * 16384;      // two images (y1, df) of 16 KB per 128-row tile, pairs of tiles
}

extern "C" int u2gnn_ffn_tc_bwd(const float* y1, const float* df, const float* dz, int64_t M, int d, int ff,
                                const void* packed, float hidden_scale, uint64_t seed, uint32_t stream_hidden, int thr,
                                float* dy1, float* dW1, float* db1, float* dW2, void* workspace, size_t workspace_bytes,
                                u2gnn_stream_t stream) {
    if (!y1 || !df || !dz || !packed || !dy1 || !dW1 || !db1 || !dW2 || !workspace || M < 0 || thr < 0 || thr > 255) return U2GNN_EINVAL;
    if (workspace_bytes < u2gnn_ffn_tc_bwd_workspace_bytes(M)) return U2GNN_EWORKSPACE;
    if (reinterpret_cast<uintptr_t>(workspace) % 128) return U2GNN_EALIGN;
    if (d < 1 || d > 64 || ff < 128 || ff % 128 || ff > 2048) return U2GNN_EUNSUPPORTED;
    if (d == 64 && ((reinterpret_cast<uintptr_t>(y1) | reinterpret_cast<uintptr_t>(df) | reinterpret_cast<uintptr_t>(dz) |
                     reinterpret_cast<uintptr_t>(dy1)) % 16))
        return U2GNN_EALIGN;
    if (M == 0) return U2GNN_OK;
    // workspace: bf16 swizzled tile images of y1 and df (written by dgrad, bulk-copied by wgrad)
    const size_t half = u2gnn_ffn_tc_bwd_workspace_bytes(M) / 2;
    uint8_t* xb = static_cast<uint8_t*>(workspace);
    uint8_t* fb = xb + half;
    int rc = ffn_tc_dgrad_launch(y1, df, dz, dy1, M, d, ff, packed, seed, stream_hidden, thr, xb, fb, as_stream(stream));
    if (rc != U2GNN_OK) return rc;
    rc = ffn_tc_wgrad_launch(xb, fb, M, d, ff, packed, hidden_scale, seed, stream_hidden, thr, dW1, db1, dW2, as_stream(stream));
    if (rc != U2GNN_OK) return rc;
    U2GNN_CHECK_LAUNCH();
}
