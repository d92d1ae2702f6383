/* u2gnn_b200 — C-ABI of the B200-native U2GNN train-step hot path.
 *
 * Drop-in boundary.  The reference has no native boundary for the model itself
 * (it is stock PyTorch eager); its only native code is the Cython-wrapped C++ sampler
 * (U2GNN_pytorch/log_uniform/log_uniform.pyx:16-40 -> Log_Uniform_Sampler.cpp).  Each entry point
 * below therefore cites the reference call it replaces.  Conventions for every function:
 *   - plain pointers + sizes, no torch types; all pointers are DEVICE pointers unless named host_*
 *   - the caller owns every buffer (inputs, outputs, workspaces); the library never allocates,
 *     never synchronises and launches only on the `stream` argument (a cudaStream_t)
 *   - returns 0 (U2GNN_OK) or a negative U2GNN_E* code; u2gnn_strerror() names it
 *   - dropout uses an explicit counter-based stream (seed, stream id) — see csrc/rng.cuh;
 *     p is quantised to thr/256, thr = 0 disables the site
 *   - index tensors are int64 (the reference dtype); feature tensors are fp32 row-major
 */
#ifndef U2GNN_B200_H
#define U2GNN_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* u2gnn_stream_t; /* cudaStream_t */

enum {
    U2GNN_OK = 0,
    U2GNN_EINVAL = -1,      /* bad shape / null pointer / unsupported size */
    U2GNN_EALIGN = -2,      /* pointer or leading dimension not aligned as required */
    U2GNN_EUNSUPPORTED = -3,/* feature size outside the kernels' range */
    U2GNN_ELAUNCH = -4,     /* CUDA launch error (cudaPeekAtLastError) */
    U2GNN_EWORKSPACE = -5,  /* workspace too small */
    U2GNN_EDEVICE = -6      /* not an sm_100 device */
};

const char* u2gnn_strerror(int code);
int u2gnn_version(void);
/* 0 when the current device can run the sm_100a kernels */
int u2gnn_device_check(void);
/* engine dropout stream: writes the keep-mask word (32 elements) for group g — host helper used
 * by tests to pin the device stream against the oracle restatement */
uint32_t u2gnn_rng_mask_word_host(uint64_t seed, uint32_t stream, uint64_t group, int thr);

/* ---- K1: row gather.  Replaces F.embedding(input_x, X_concat) / F.embedding(input_x, output_Tr)
 *      (pytorch_U2GNN_Sup.py:32,39).  out[i,:] = table[idx[i],:]; idx_stride lets the caller
 *      gather only column 0 of input_x (idx element i is idx[i*idx_stride]).  Bit-exact copy.
 *      An index outside [0, n_table) - where F.embedding raises - yields a ZERO row and sets bit 1 of the device
 *      error word err (may be NULL: the row-sharded sampled-row gather uses the zero rows on purpose). */
int u2gnn_gather_rows(const float* table, int64_t n_table, int d, const int64_t* idx, int64_t n_idx,
                      int64_t idx_stride, float* out, int* err, u2gnn_stream_t stream);
/* backward of the re-gather (autograd of pytorch_U2GNN_Sup.py:39): dst[idx[i],:] += grad[i,:].
 * dst must be initialised by the caller.  deterministic=0 uses fp32 atomics. */
int u2gnn_scatter_add_rows(const float* grad, int64_t n_idx, int d, const int64_t* idx, int64_t idx_stride,
                           float* dst, int64_t n_dst, u2gnn_stream_t stream);
/* deterministic variant: CSR transpose of the index list (built once per batch) + ordered sums */
size_t u2gnn_index_transpose_workspace_bytes(int64_t n_idx, int64_t n_dst);
int u2gnn_index_transpose_build(const int64_t* idx, int64_t n_idx, int64_t idx_stride, int64_t n_dst,
                                int64_t* t_rowptr /*[n_dst+1]*/, int64_t* t_pos /*[n_idx]*/,
                                void* workspace, size_t workspace_bytes, u2gnn_stream_t stream);
int u2gnn_scatter_add_rows_det(const float* grad, int d, const int64_t* t_rowptr, const int64_t* t_pos,
                               float* dst, int64_t n_dst, int accumulate, u2gnn_stream_t stream);

/* ---- K4: graph sum-pooling.  Replaces torch.spmm(graph_pool, output_Tr) (pytorch_U2GNN_Sup.py:41)
 *      with the CSR form of the reference's COO operator (train_pytorch_U2GNN_Sup.py:73-89):
 *      out[g,:] = sum_{n in [rowptr[g], rowptr[g+1])} x[n,:], summed in ascending node order. */
int u2gnn_rowptr_from_coo(const int64_t* coo_rows, int64_t nnz, int64_t num_graphs, int64_t* rowptr,
                          u2gnn_stream_t stream);
int u2gnn_segment_sum(const float* x, int64_t n, int d, const int64_t* rowptr, int64_t num_graphs, float* out,
                      u2gnn_stream_t stream);
int u2gnn_segment_sum_bwd(const float* grad_out, int64_t num_graphs, int d, const int64_t* rowptr, float* grad_x,
                          int64_t n, int accumulate, u2gnn_stream_t stream);

/* ---- fp32 building blocks of the encoder layer (nn.TransformerEncoderLayer, post-norm, 1 head;
 *      pytorch_U2GNN_Sup.py:20-21,35 -> torch/nn/modules/transformer.py:944-982) ---- */
/* C[M,N] = epi(alpha * op(A) op(B) + bias) (+ beta*C).  op(A) is M x K: A[m*lda+k] (ta=0) or
 * A[k*lda+m] (ta=1); op(B) is K x N: B[k*ldb+n] (tb=0) or B[n*ldb+k] (tb=1).
 * epi flags: 1 = add bias[n]; 2 = ReLU; 4 = dropout with (seed, stream, thr) on linear index (rng_row0+m)*N+n;
 *            8 = multiply by scale*(aux[m*ldaux+n] > 0) (ReLU/dropout backward through saved output);
 *            16 = accumulate into C with atomics (split-K along K over `splitk` slices). */
int u2gnn_sgemm(int ta, int tb, int64_t M, int N, int64_t K, float alpha, const float* A, int64_t lda,
                const float* B, int64_t ldb, float beta, float* C, int64_t ldc, const float* bias, int epi,
                uint64_t seed, uint32_t rng_stream, int thr, int64_t rng_row0, const float* aux, int64_t ldaux,
                float aux_scale, int splitk, u2gnn_stream_t stream);
/* The same products ON THE TENSOR CORES at fp32 accuracy (csrc/gemm_split.cu): every fp32 operand value is staged as
 * hi = bf16(a), lo = bf16(a - hi) and every product evaluated as A_hi B_hi + A_lo B_hi + A_hi B_lo (three tcgen05.mma per
 * k-step, fp32 accumulation in tensor memory; relative error of a product <= ~2^-16).  precision="fp32" runs F.linear and its
 * autograd (linear1 / linear2, in_proj / out_proj of torch/nn/modules/transformer.py:944-982) through these two entry points.
 *   rows : C[M,N] = epi(A[M,K] op(W) + bias) (+ beta*C);  W[n*ldw+k] (w_kn=0) or W[k*ldw+n] (w_kn=1); epi bits 1 / 2 / 4 / 8 as
 *          u2gnn_sgemm (dropout on linear index (rng_row0+m)*N+n).  Any M, K, N, lda, ldw, ldc; 16-byte aligned shapes use
 *          128-bit accesses.
 *   wgrad: dW[n1*ldw_n1 + n2*ldw_n2] += sum_m A[m*lda+n1] B[m*ldb+n2];  db[n1] += sum_m A[m*lda+n1] (db may be NULL).
 *          Accumulates with atomics (the destination strides make a transposed gradient a stride swap). */
int u2gnn_gemm_split_rows(const float* A, int64_t M, int K, int64_t lda, const float* W, int w_kn, int64_t ldw, int N,
                          const float* bias, int epi, uint64_t seed, uint32_t rng_stream, int thr, int64_t rng_row0,
                          const float* aux, int64_t ldaux, float aux_scale, float beta, float* C, int64_t ldc,
                          const void* packed_w, u2gnn_stream_t stream);
/* packed_w (optional, both rows entry points): the weights already converted to the kernel's swizzled bf16 operand images by
 * u2gnn_gemm_split_pack (u2gnn_gemm_split_packed_bytes(N, K) bytes, 16-byte aligned; the hi / lo images of the split mode serve the plain
 * mode too).  A step's weights are then ONE asynchronous bulk copy instead of an L2 read + conversion by all 256 threads: worth one extra
 * tiny launch per weight for batches of more than a few thousand rows.  W must still be passed (shapes, and when packed_w is NULL). */
size_t u2gnn_gemm_split_packed_bytes(int N, int K);
int u2gnn_gemm_split_pack(const float* W, int w_kn, int64_t ldw, int N, int K, void* packed, size_t packed_size, u2gnn_stream_t stream);
int u2gnn_gemm_split_wgrad(const float* A, int64_t M, int N1, int64_t lda, const float* B, int N2, int64_t ldb, float* dW,
                           int64_t ldw_n1, int64_t ldw_n2, float* db, u2gnn_stream_t stream);
/* The rows kernel above with ONE bf16 product per k-step (plain bf16 mode): A is bf16 row-major (lda, K multiples of 8, 16-byte aligned),
 * W fp32 as above, the result fp32 (c_bf16 = 0) or bf16 (c_bf16 = 1: beta must be 0), the aux mask fp32 or bf16 (only its sign is
 * used).  Loops over K inside the kernel (any K) and over 128-column slices of N, with the u2gnn_sgemm epilogue bits fused: the
 * bf16 FFN of 64 < d <= 128 (csrc/ffn_wide.cu, BASELINE.json configs[2]) runs linear1 + ReLU + dropout, linear2, the hidden's gradient
 * with its mask and dy1 as one launch each.  bf16 results / masks need 4-element aligned shapes (N, ldc, ldaux multiples of 4). */
int u2gnn_gemm_tc_rows_kloop(const void* A, int64_t M, int K, int64_t lda, const float* W, int w_kn, int64_t ldw, int N,
                             const float* bias, int epi, uint64_t seed, uint32_t rng_stream, int thr, int64_t rng_row0,
                             const void* aux, int aux_bf16, int64_t ldaux, float aux_scale, float beta, void* C, int c_bf16,
                             int64_t ldc, const void* packed_w, u2gnn_stream_t stream);
/* out[n] (+)= sum_m A[m*lda+n] */
int u2gnn_colsum(const float* A, int64_t M, int N, int64_t lda, float* out, int accumulate, u2gnn_stream_t stream);

/* short-sequence self-attention (attn_axis="neighbors", S <= 32): one sequence per node.
 * qkv[B, S, 3d] (q | k | v per row), ctx[B, Sq, d]; Sq = S, or 1 for the dead-row-eliminated last
 * timestep (only query position 0; then q rows are read from qkv position 0).  softmax(q k^T/sqrt(d))
 * with dropout on the probabilities (linear index (b*Sq+i)*S+j), then @ v. */
int u2gnn_seqattn_fwd(const float* qkv, int64_t B, int S, int Sq, int d, uint64_t seed, uint32_t rng_stream, int thr,
                      float* ctx, u2gnn_stream_t stream);
/* dqkv[B, S, 3d] is fully written (rows without a query get dq = 0) */
int u2gnn_seqattn_bwd(const float* qkv, const float* dctx, int64_t B, int S, int Sq, int d, uint64_t seed,
                      uint32_t rng_stream, int thr, float* dqkv, u2gnn_stream_t stream);
/* long-sequence pieces (attn_axis="nodes"): row softmax of scores[M, N] in place with dropout, and
 * its backward  ds = p * (dp*mask - rowsum(dp*mask*p)) given the un-dropped probabilities p. */
int u2gnn_softmax_rows_fwd(float* scores, int64_t M, int64_t N, float* probs_dropped, uint64_t seed,
                           uint32_t rng_stream, int thr, u2gnn_stream_t stream);
int u2gnn_softmax_rows_bwd(const float* probs, float* dprobs_inout, int64_t M, int64_t N, uint64_t seed,
                           uint32_t rng_stream, int thr, u2gnn_stream_t stream);

/* z = res + dropout(a);  y = LayerNorm(z) * gamma + beta  (eps 1e-5, biased variance).
 * res may be null (plain LayerNorm of a).  Saves z (pre-norm) and stats[M,2] = (mean, rstd). */
int u2gnn_add_dropout_ln_fwd(const float* res, const float* a, int64_t M, int d, uint64_t seed, uint32_t rng_stream,
                             int thr, const float* gamma, const float* beta, float* z, float* y, float* stats,
                             u2gnn_stream_t stream);
/* dz = LN backward of dy; dgamma/dbeta accumulated (atomics); da = dz * dropout mask (may be null). */
int u2gnn_add_dropout_ln_bwd(const float* dy, const float* z, const float* stats, int64_t M, int d,
                             const float* gamma, uint64_t seed, uint32_t rng_stream, int thr, float* dz, float* da,
                             float* dgamma, float* dbeta, u2gnn_stream_t stream);
/* same, with da optionally stored as bf16 (da_bf16 = 1: its consumers are tensor-core kernels that round on load, so
 * rounding once here is bit-identical at half the bytes) and dasum[d] += colsum(dz * dropout mask) (fp32 values before
 * rounding; may be null; works without da) - the bias gradient of the linear layer that produced a.  Both options need d in {4, 8, 16, 32, 64, 128}.
 * da_bf16 = 2 (d = 64): da is written as bf16 swizzled [128 x 64] tile images (row r -> image r / 128, 128-byte rows, 16-byte chunk
 * index XOR (r & 7)), the operand format u2gnn_ffn_tc_bwd bulk-copies (df_img): the fp32 gradient never reaches HBM. */
int u2gnn_add_dropout_ln_bwd_ex(const float* dy, const float* z, const float* stats, int64_t M, int d,
                                const float* gamma, uint64_t seed, uint32_t rng_stream, int thr, float* dz, void* da,
                                int da_bf16, float* dgamma, float* dbeta, float* dasum, u2gnn_stream_t stream);
/* y = LayerNorm(z) from saved stats (re-materialises a layer input from its pre-norm value) */
int u2gnn_ln_apply(const float* z, const float* stats, int64_t M, int d, const float* gamma, const float* beta,
                   float* y, u2gnn_stream_t stream);
/* elementwise helpers: y = x * keepmask * scale ; y += x */
int u2gnn_dropout_apply(const float* x, int64_t numel, uint64_t seed, uint32_t rng_stream, int thr, float* y,
                        u2gnn_stream_t stream);
int u2gnn_axpy(float alpha, const float* x, float* y, int64_t numel, u2gnn_stream_t stream);
/* strided row copy: dst[i*ld_dst + c] = src[i*ld_src + c], c < d (position-0 select, concat) */
int u2gnn_copy_rows(const float* src, int64_t ld_src, float* dst, int64_t ld_dst, int64_t rows, int d, int accumulate,
                    u2gnn_stream_t stream);

/* ---- K5: classifier head + label-smoothed soft cross-entropy.
 *      Replaces pytorch_U2GNN_Sup.py:42-44,48-59 and train_pytorch_U2GNN_Sup.py:140-142.
 *      scores[G,C] += dropout(ge[G,d]) @ W[C,d]^T + b   (called once per U2GNN layer)        */
int u2gnn_head_fwd(const float* ge, int64_t G, int d, const float* W, const float* b, int C, uint64_t seed,
                   uint32_t rng_stream, int thr, float* scores, int accumulate, u2gnn_stream_t stream);
/* loss = mean_g sum_c -t[g,c] log_softmax(scores)[g,c], t = label_smoothing(labels, C, smoothing);
 * writes loss[1] and dscores[G,C] (gradient of the mean loss; G_total lets data-parallel ranks
 * divide by the global number of graphs). */
int u2gnn_soft_ce_fwd_bwd(const float* scores, const int64_t* labels, int64_t G, int C, float smoothing,
                          int64_t G_total, float* loss, float* dscores, u2gnn_stream_t stream);
/* dW[C,d] += dscores^T @ dropout(ge); db[C] += colsum(dscores); dge[G,d] = (dscores @ W) * mask */
int u2gnn_head_bwd(const float* dscores, const float* ge, int64_t G, int d, const float* W, int C, uint64_t seed,
                   uint32_t rng_stream, int thr, float* dW, float* db, float* dge, u2gnn_stream_t stream);

/* ---- K6: log-uniform candidate sampler.  Replaces LogUniformSampler.sample
 *      (log_uniform.pyx:29-34 -> Log_Uniform_Sampler.cpp:57-71, std::default_random_engine(1111)).
 *      Draw i of the reference's sequential stream is recomputed independently on the device by
 *      LCG skip-ahead; the first `size` distinct ids in stream order are kept.  state_inout[0] is
 *      the minstd_rand0 state (1111 initially) and is advanced past the consumed draws;
 *      out_ids[size] holds the ids in first-occurrence order, out_tries[0] the reference's
 *      num_tries.  Single-CTA kernel; workspace from u2gnn_logu_sample_workspace_bytes. */
size_t u2gnn_logu_sample_workspace_bytes(int64_t size);
int u2gnn_logu_sample(int64_t range_max, int64_t size, uint32_t* state_inout, int64_t* out_ids, int32_t* out_tries,
                      void* workspace, size_t workspace_bytes, u2gnn_stream_t stream);
/* sample_unique (Log_Uniform_Sampler.cpp:73-88; binding log_uniform.pyx:25-27): `size` distinct ids NOT in labels[n_labels];
 * draws that hit a label consume the engine stream and are dropped.  Same id set and advanced state as the reference.
 * Requires size + n_labels <= range_max (otherwise the reference may never terminate). */
size_t u2gnn_logu_sample_unique_workspace_bytes(int64_t range_max, int64_t size);
int u2gnn_logu_sample_unique(int64_t range_max, int64_t size, const int64_t* labels, int64_t n_labels, uint32_t* state_inout,
                             int64_t* out_ids, void* workspace, size_t workspace_bytes, u2gnn_stream_t stream);
/* expected_count (Log_Uniform_Sampler.cpp:23-32): out[i] = -expm1(tries * log1p(-prob[ids[i]])) */
int u2gnn_logu_expected_count(int64_t range_max, const int32_t* tries, const int64_t* ids, int64_t n, float* out,
                              u2gnn_stream_t stream);

/* ---- K7: fused sampled softmax.  Replaces SampledSoftmax.sampled (sampled_softmax.py:36-56):
 *      loss[i] = -log( exp(x_i . W[y_i]) / sum_s exp(x_i . W[ids[s]]) ), no max-subtraction, no
 *      log-Q correction.  Saves denom[N] for the backward.
 *      Backward: dx[N,D] and a DENSE dW[V,D] accumulation (the reference's W.grad is dense).
 *      samp_rows (may be NULL): the ns sampled rows already gathered into one [ns, D] buffer (row-sharded table: the
 *      owners' rows all-reduced, parallel.py) - ids are then not dereferenced into W, and the backward adds the sampled
 *      rows' gradient into dsamp[ns, D] instead of dW[ids] (samp_rows and dsamp go together).
 *      err (may be NULL): device error word; bit 0 is set when a label lies outside [0, V) - the reference's
 *      index_select raises there (sampled_softmax.py:45); the kernels give such a row loss 0 and no gradient. */
int u2gnn_sampled_softmax_fwd(const float* x, const int64_t* labels, int64_t N, int D, const float* W, int64_t V,
                              const int64_t* ids, int ns, const float* samp_rows, float* loss, float* denom, int* err,
                              u2gnn_stream_t stream);
int u2gnn_sampled_softmax_bwd(const float* dloss, const float* x, const int64_t* labels, int64_t N, int D,
                              const float* W, int64_t V, const int64_t* ids, int ns, const float* samp_rows,
                              const float* denom, float* dx, float* dW, float* dsamp, int* err, u2gnn_stream_t stream);
/* TF-model variant of the loss (U2GNN_tf/model_U2GNN_Unsup_multi.py:54-58, tf.nn.sampled_softmax_loss with
 * its defaults): t_i = x_i.W[y_i] + bias[y_i] - log true_q[i], l_is = x_i.W[s] + bias[s] - log samp_q[s] (accidental hits
 * s == y_i removed), loss_i = log(exp t_i + sum_s exp l_is) - t_i.  true_q / samp_q are the expected counts of the labels and of
 * the sampled ids (u2gnn_logu_expected_count, Log_Uniform_Sampler.cpp:23-32).  denom[N] is saved for the backward, which also
 * accumulates the bias gradient dbias[V] (+=, like dW). */
int u2gnn_sampled_softmax_tf_fwd(const float* x, const int64_t* labels, int64_t N, int D, const float* W, const float* bias,
                                 int64_t V, const int64_t* ids, int ns, const float* true_q, const float* samp_q, float* loss,
                                 float* denom, int* err, u2gnn_stream_t stream);
int u2gnn_sampled_softmax_tf_bwd(const float* dloss, const float* x, const int64_t* labels, int64_t N, int D, const float* W,
                                 const float* bias, int64_t V, const int64_t* ids, int ns, const float* true_q,
                                 const float* samp_q, const float* denom, float* dx, float* dW, float* dbias,
                                 int* err, u2gnn_stream_t stream);

/* ---- K8: fused global-norm clip + Adam.  Replaces clip_grad_norm_(params, 0.5) + Adam.step()
 *      (train_pytorch_U2GNN_Sup.py:145,160-161) over ONE flat parameter arena.
 *      sqnorm: sumsq[0] += sum(g^2) (caller zeroes it; all-reduced across ranks by the host side).
 *      adam: coef = min(1, max_norm/(sqrt(sumsq)+1e-6)); g *= coef; standard Adam, step 1-based. */
int u2gnn_grad_sqnorm(const float* g, int64_t n, float* sumsq, u2gnn_stream_t stream);
int u2gnn_clip_adam(float* p, const float* g, float* m, float* v, int64_t n, const float* sumsq, float max_norm,
                    float lr, float beta1, float beta2, float eps, int64_t step, u2gnn_stream_t stream);

/* ---- fused bf16 tensor-core FFN block (tcgen05 / TMEM), the path's dense contraction.
 *      Replaces linear1 -> ReLU -> dropout -> linear2 -> dropout -> residual -> norm2 of
 *      nn.TransformerEncoderLayer (torch/nn/modules/transformer.py:950-958,977-982); the [M, ff] hidden
 *      never leaves the SM.  d <= 64, ff a multiple of 128.  Weights are packed once per optimiser step
 *      into pre-swizzled bf16 shared-memory images (hidden_scale = 1/(1-p) of the hidden dropout is
 *      folded into the W2 image).
 *      z[M,d] = y1 + dropout_out( dropout_hidden(relu(y1 W1^T + b1)) W2^T + b2 );
 *      stats[M,2] = (mean, rstd) of z;  xnext[M,d] = LayerNorm(z)*gamma + beta (may be null).
 *      mask_out (may be NULL): u2gnn_ffn_tc_mask_bytes(M, ff) bytes receiving one bit per hidden activation (ReLU live AND
 *      dropout keep), which u2gnn_ffn_tc_bwd(fwd_mask) consumes instead of recomputing the sign of the hidden and the dropout
 *      stream. */
size_t u2gnn_ffn_tc_packed_bytes(int d, int ff);
int u2gnn_ffn_tc_prepare(const float* W1, const float* b1, const float* W2, const float* b2, int d, int ff,
                         float hidden_scale, void* packed, size_t packed_size, u2gnn_stream_t stream);
size_t u2gnn_ffn_tc_mask_bytes(int64_t M, int ff);
int u2gnn_ffn_tc_fwd(const float* y1, int64_t M, int d, int ff, const void* packed, uint64_t seed,
                     uint32_t stream_hidden, uint32_t stream_out, int thr, const float* gamma, const float* beta,
                     float* z, float* stats, float* xnext, void* mask_out, u2gnn_stream_t stream);

/* backward of the block above (autograd of linear1/ReLU/dropout/linear2, transformer.py:977-982) with the
 * hidden recomputed on chip: dy1[M,d] = dz + dPre W1 (dz = gradient at the residual sum, df = gradient at
 * the linear2 output after the output dropout; dy1 may alias dz); dW1[ff,d], db1[ff], dW2[d,ff] are
 * ACCUMULATED (atomics).  db2 = colsum(df) is left to u2gnn_colsum.  workspace (128-byte aligned) holds the bf16 tile
 * images of y1 / df both tensor-core kernels read and the 1-bit ReLU-and-keep mask words the weight-gradient kernel hands
 * to the input-gradient kernel (so the hidden is recomputed once, not twice).
 * y1_img / df_img (may be NULL): the same two operands ALREADY stored as bf16 swizzled [128 x 64] tile images of
 * u2gnn_ffn_tc_image_bytes(M) bytes (128-byte aligned; rows >= M of the last tile zero) - what u2gnn_gemm_tc_rows_ln (y_img)
 * and u2gnn_add_dropout_ln_bwd_ex (da_bf16 = 2) write; the fp32 pointer of an operand given as an image may be NULL.
 * fwd_mask (may be NULL): the mask words u2gnn_ffn_tc_fwd(mask_out) wrote for the same rows, weights and dropout stream: the
 * weight-gradient kernel then loads them instead of evaluating the dropout stream and extracting the sign of the hidden. */
size_t u2gnn_ffn_tc_bwd_workspace_bytes(int64_t M);
size_t u2gnn_ffn_tc_image_bytes(int64_t M);
int u2gnn_ffn_tc_bwd(const float* y1, const float* df, const void* y1_img, const void* df_img, const void* fwd_mask, const float* dz,
                     int64_t M, int d, int ff, const void* packed, float hidden_scale, uint64_t seed, uint32_t stream_hidden, int thr,
                     float* dy1, float* dW1, float* db1, float* dW2, void* workspace, size_t workspace_bytes, u2gnn_stream_t stream);

/* ---- bf16 FFN for 64 < d <= 128 (csrc/ffn_wide.cu; BASELINE.json configs[2], d = 65): the FFN runs as the general tcgen05 GEMMs
 *      below with the [M, ff] hidden materialised in bf16; these are the elementwise steps between them (ReLU + dropout of
 *      transformer.py:977-982 and their autograd).  fwd: h <- relu(h) * keep * scale in place (keep = engine dropout stream over
 *      element row * ff + col; ff a multiple of 32);  bwd: dh <- (h > 0) ? dh * scale : 0 with h the saved forward result. */
int u2gnn_relu_dropout_bf16(void* h, int64_t M, int ff, uint64_t seed, uint32_t rng_stream, int thr, float scale, u2gnn_stream_t stream);
/* fp32 rows [M, d] -> bf16 rows [M, dp] zero-padded (dp a multiple of 8): the 16-byte-aligned operand copy of y1 / dF the GEMMs stream */
int u2gnn_pad_rows_bf16(const float* src, int64_t M, int d, void* dst, int dp, u2gnn_stream_t stream);
int u2gnn_relu_dropout_bwd_bf16(void* dh, const void* h, int64_t M, int ff, float scale, u2gnn_stream_t stream);

/* ---- bf16 tensor-core GEMMs for the attention-block projections of the bf16 mode (csrc/gemm_tc.cu): the
 *      F.linear calls inside nn.MultiheadAttention (in_proj / out_proj) and their autograd.
 *      rows : C[M,N] = A[M,K] W^T (+ bias) (+ beta*C);  W is [N,K] (w_kn = 0) or [K,N] (w_kn = 1); K, N <= 256
 *      wgrad: dW[N1,N2] += A[M,N1]^T B[M,N2];  db[N1] += colsum(A) (db may be null);  N1 <= 256, N2 <= 64 */
int u2gnn_gemm_tc_rows(const float* A, int64_t M, int K, int64_t lda, const float* W, int w_kn, int N,
                       const float* bias, float beta, float* C, int64_t ldc, u2gnn_stream_t stream);
int u2gnn_gemm_tc_wgrad(const float* A, int64_t M, int N1, int64_t lda, const float* B, int N2, int64_t ldb,
                        float* dW, float* db, u2gnn_stream_t stream);

/* ---- short-sequence attention core on the tensor cores (bf16 mode; d = 64, 2 <= S <= 32, every query row live).
 *      Same contract as u2gnn_seqattn_fwd / _bwd with Sq = S (qkv[B,S,192] -> ctx[B*S,64]; dctx -> dqkv[B*S,192]). */
int u2gnn_seqattn_tc_fwd(const float* qkv, int64_t B, int S, int d, uint64_t seed, uint32_t rng_stream, int thr,
                         float* ctx, u2gnn_stream_t stream);
int u2gnn_seqattn_tc_bwd(const float* qkv, const float* dctx, int64_t B, int S, int d, uint64_t seed,
                         uint32_t rng_stream, int thr, float* dqkv, u2gnn_stream_t stream);

/* ---- bf16 activation I/O for the attention block of the bf16 mode.  Every consumer of qkv / ctx / dctx / dqkv rounds them
        to bf16 before its tensor-core product, so the producers store them as bf16 row-major (rounded once, bit-identical
        results) and the block moves half the HBM bytes.  *_bf16 flags: 0 = fp32 buffer, 1 = bf16 buffer; leading
        dimensions are in ELEMENTS; beta must be 0 for a bf16 C. ---- */
int u2gnn_gemm_tc_rows_ex(const void* A, int a_bf16, int64_t M, int K, int64_t lda, const float* W, int w_kn, int N,
                          const float* bias, float beta, void* C, int c_bf16, int64_t ldc, u2gnn_stream_t stream);
/* FUSED GATHER (SURVEY.md 8(a) a3 "fused: no write"): where an entry point below takes an index array next to a row operand
 * (b_idx / res_idx / x_idx, NULL = rows in place), row r of that operand is row idx[r] of a TABLE with *_rows rows - the
 * F.embedding(input_x, X_concat) of pytorch_U2GNN_Sup.py:32 evaluated inside its consumers, so that the gathered [N, k+1, d]
 * tensor of the first timestep is never written to HBM.  Indexed operands are fp32 with 64 columns; an index outside the table
 * reads as a zero row (u2gnn_inproj_seqattn_tc_fwd also sets bit 1 of the device error word, as u2gnn_gather_rows does). */
int u2gnn_gemm_tc_wgrad_ex(const void* A, int a_bf16, int64_t M, int N1, int64_t lda, const void* B, int b_bf16, int N2,
                           int64_t ldb, const int64_t* b_idx, int64_t b_rows, float* dW, float* db, u2gnn_stream_t stream);
/* backward of one attention-block projection in a single pass over its output gradient A[M,N1] (bf16; N1 = 64, 128, 192
 * or 256): input gradient C[M,64] = A W (+ beta*C; W = the layer's weight [N1,64]; C fp32, or bf16 with beta 0) AND
 * dW[N1,64] += A^T B, db[N1] += colsum(A) (B[M,64] = the layer's input rows, fp32 or bf16; db may be null).  Same results as
 * u2gnn_gemm_tc_rows_ex(w_kn = 1) + u2gnn_gemm_tc_wgrad_ex on the same operands.  (autograd of the in_proj / out_proj F.linear
 * calls of nn.MultiheadAttention, which pytorch_U2GNN_Sup.py:20-21 instantiates through nn.TransformerEncoderLayer) */
int u2gnn_gemm_tc_dgrad_wgrad(const void* A, int64_t M, int N1, int64_t lda, const void* B, int b_bf16, int64_t ldb,
                              const int64_t* b_idx, int64_t b_rows, const float* W, void* C, int c_bf16, int64_t ldc, float beta,
                              float* dW, float* db, u2gnn_stream_t stream);
/* out_proj + dropout + residual + LayerNorm1 in one kernel (torch/nn/modules/transformer.py:946,969-972:
 * x = norm1(x + dropout1(self_attn(x)))): z[M,64] = res + dropout(A[M,K] W^T + bias), y = LayerNorm(z) * gamma + beta,
 * stats[M,2] = (mean, rstd).  N = d = 64 only; res row stride ldres (elements) so the last timestep can read position 0 of
 * each sequence in place.  Same arithmetic as u2gnn_gemm_tc_rows_ex followed by u2gnn_add_dropout_ln_fwd (bit-identical).
 * y_img (may be NULL): y additionally as bf16 swizzled tile images (u2gnn_ffn_tc_image_bytes(M) bytes) for u2gnn_ffn_tc_bwd. */
int u2gnn_gemm_tc_rows_ln(const void* A, int a_bf16, int64_t M, int K, int64_t lda, const float* W, int w_kn,
                          const float* bias, const float* res, int64_t ldres, const int64_t* res_idx, int64_t res_rows,
                          uint64_t seed, uint32_t rng_stream, int thr, const float* gamma, const float* beta, float* z, float* y,
                          float* stats, void* y_img, u2gnn_stream_t stream);
int u2gnn_seqattn_tc_fwd_ex(const void* qkv, int64_t B, int S, int d, uint64_t seed, uint32_t rng_stream, int thr,
                            void* ctx, int io_bf16, u2gnn_stream_t stream);
int u2gnn_seqattn_tc_bwd_ex(const void* qkv, const void* dctx, int64_t B, int S, int d, uint64_t seed,
                            uint32_t rng_stream, int thr, void* dqkv, int io_bf16, u2gnn_stream_t stream);

/* in_proj + attention core of one timestep in ONE kernel (bf16 mode, d = 64, every query row live): qkv = x W_in^T + b_in
 * (nn.MultiheadAttention in_proj: w_in [192,64], b_in [192]) is computed per tile, written once as bf16 qkv_out[B*S,192] for the
 * backward, and consumed from shared memory; ctx[B*S,64] bf16.  Same results as u2gnn_gemm_tc_rows_ex (bf16 out) followed by
 * u2gnn_seqattn_tc_fwd_ex(io_bf16 = 1). */
int u2gnn_inproj_seqattn_tc_fwd(const float* x, const int64_t* x_idx, int64_t x_rows, int64_t B, int S, int d, const float* w_in,
                                const float* b_in, uint64_t seed, uint32_t rng_stream, int thr, void* qkv_out, void* ctx, int* err,
                                u2gnn_stream_t stream);
/* last timestep of a U2GNN layer (only query position 0 of each node live): same contract as
 * u2gnn_seqattn_fwd / _bwd with Sq = 1, with qkv (and dqkv) optionally stored as bf16 [B*S, 3d]; ctx / dctx are fp32 [B, d].
 * d in {32, 64}, 2 <= S <= 32. */
int u2gnn_seqattn_last_fwd_ex(const void* qkv, int qkv_bf16, int64_t B, int S, int d, uint64_t seed, uint32_t rng_stream,
                              int thr, float* ctx, u2gnn_stream_t stream);
int u2gnn_seqattn_last_bwd_ex(const void* qkv, const float* dctx, int io_bf16, int64_t B, int S, int d, uint64_t seed,
                              uint32_t rng_stream, int thr, void* dqkv, u2gnn_stream_t stream);

/* ---- device-side batch builder (replaces the host loop of get_batch_data,
        train_pytorch_U2GNN_Sup.py:91-119 / train_pytorch_U2GNN_UnSup.py:96-128).
        Dataset adjacency as one CSR over dataset-wide node ids (g_rowptr[V+1], g_col[E]); the batch is n_graphs selected
        graphs: graph_start[g] = first dataset-wide node id of graph g, batch_off[n_graphs+1] = prefix sums of their node
        counts (= the pooling rowptr), n_nodes = batch_off[n_graphs].  Writes input_x[n_nodes, k+1] (batch-local ids:
        the node itself, then k neighbours drawn uniformly with replacement; isolated nodes repeat themselves) and, if
        non-NULL, node_global[n_nodes] (dataset-wide ids: the X_concat gather list / input_y).  The draw is a pure
        function of (seed, rng_stream, node, slot). ---- */
int u2gnn_build_batch(const int64_t* g_rowptr, const int64_t* g_col, const int64_t* graph_start,
                      const int64_t* batch_off, int64_t n_graphs, int64_t n_nodes, int k, uint64_t seed,
                      uint32_t rng_stream, int64_t* input_x, int64_t* node_global, u2gnn_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* U2GNN_B200_H */
