/* u2gnn_b200_probe.h - entry points of the PROBE library (libu2gnn_b200_probe.so), not of the product.
 *
 * The probe library is the product library's sources compiled with -DU2GNN_PROBE_BUILD (which adds the clock-stamp tracing
 * instantiations of the FFN kernels) plus csrc/probe/*.cu: hardware layout self-tests and micro-benchmarks.  None of this
 * has a reference counterpart; tools/ and the layout self-tests in tests/ load it, the product path never does.
 * libu2gnn_b200.so exports none of these symbols and keeps no process-global state. */
#ifndef U2GNN_B200_PROBE_H
#define U2GNN_B200_PROBE_H
#include "u2gnn_b200.h"
#ifdef __cplusplus
extern "C" {
#endif

/* clock-stamp tracing of CTA 0 of the next FFN forward / backward launches (tools/trace_ffn.py, tools/trace_ffn_bwd.py):
 * buf = device buffer of 64 x 1024 uint32 (slot 0 = MMA warp, 1.. = epilogue warps; the dgrad kernel uses slots 32..); NULL = off.
 * PROCESS-GLOBAL state - the reason it lives here and not in the product library. */
int u2gnn_ffn_tc_set_trace(void* buf);

/* ---- tcgen05 plumbing self-test (csrc/probe/tc_selftest.cu): one CTA runs a [128 x N x K] bf16 GEMM through
 *      each operand path the fused kernels use (mode 0 K-major smem, 1 MN-major smem, 2 A in tensor
 *      memory, 3 bulk-copied pre-swizzled B).  No reference counterpart: it pins hardware layout
 *      assumptions.  A, B fp32 inputs (rounded to bf16 inside), C[128, N] fp32; scratch >= 32 KB. */
int u2gnn_tc_selftest(int mode, const float* A, const float* B, float* C, int K, int N, void* scratch,
                      u2gnn_stream_t stream);

/* tcgen05.mma rate probe (tools/probe_mma.py): out[0] = issue cycles, out[1] = issue+execute cycles of `count` MMAs */
int u2gnn_tc_probe(int N, int ts, int rotate, int count, long long* out, u2gnn_stream_t stream);
/* the same for cta_group::2 (a cluster of two CTAs, M = 256; tools/probe_mma2.py): N a multiple of 32, ts 0 = SS, 1 = A in tensor memory */
int u2gnn_tc_probe2(int N, int ts, int count, long long* out, u2gnn_stream_t stream);
/* L2 reduction throughput probe (tools/probe_red.py): `groups` CTAs add into the same 32 KB tile, tile after tile;
   mode 0 red.v4.f32, 1 scalar atomicAdd, 2 plain stores, 3 red.v4.f32 thread-per-row */
int u2gnn_red_probe(float* buf, int64_t n_tiles, int groups, int mode, u2gnn_stream_t stream);
/* TMEM -> register bandwidth probe: out[0] = cycles, out[1] = bytes */
int u2gnn_tmem_bw_probe(int warps, int iters, int batch, long long* out, u2gnn_stream_t stream);

/* host evaluation of the bit arithmetic of the FFN chunk epilogue (csrc/ffn_epi.cuh; tests/test_cabi_host.py): kp16[j] = the bf16
 * factor pair (0x4000 = 2.0 kept, 0 dropped, per half) of register pair j for a natural-order keep word (epi::keep_factors16);
 * pos32[e] = epi::flag_pos(e), elem32[b] = epi::flag_elem(b): the bit order of the mask words the FFN kernels exchange. */
int u2gnn_epi_bits_host(uint32_t keep, uint32_t* kp16, int* pos32, int* elem32);

#ifdef __cplusplus
}
#endif
#endif /* U2GNN_B200_PROBE_H */
