/*
 * gatk.h -- C ABI of libgatk.so: the B200 (sm_100a) GAT-layer kernels.
 *
 * Drop-in boundary for the hot path of ArielleRosinski/pyGAT (reference files under
 * /root/reference).  Every entry point takes plain device pointers, sizes and a
 * cudaStream_t (passed as void*); no torch types.  All buffers are allocated by the
 * caller (PyTorch on the host side), the library keeps no reference after return.
 * Calls are stream-ordered and never synchronise the device.
 *
 * Return value: 0 on success, non-zero on error; gatk_last_error() returns the
 * message of the last failure on the calling thread.
 *
 * Common layout conventions
 *   N        nodes (rows = destination i, columns = source j;  layers.py:141-150)
 *   E        stored entries of the adjacency pattern, in adj.nonzero() order (layers.py:129)
 *   H        heads of one layer (models.py:17-27 builds them as separate modules)
 *   Dp       per-head width padded to 4 * 2^k floats (D = out_features, layers.py:15)
 *   "rows"   fp32 [N, H*Dp] with an explicit leading dimension (floats), head h at
 *            columns [h*Dp, (h+1)*Dp) -- this IS torch.cat(heads, dim=1) (models.py:32)
 *   rowptr   int64 [N+1];  col / trow / perm  int32 [E]  (E < 2^31 per graph shard)
 */
#ifndef GATK_H
#define GATK_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GATK_VERSION 100
#if defined(__GNUC__)
#define GATK_API __attribute__((visibility("default")))
#else
#define GATK_API
#endif

GATK_API int gatk_version(void);
GATK_API const char* gatk_last_error(void);
/* Kernels launched by the library in this process so far (every launch is counted). */
GATK_API int64_t gatk_launch_count(void);
/* Number of SMs of the current device (grid sizing), <0 on error. */
GATK_API int gatk_sm_count(void);

/* ------------------------------------------------------------------ K0: graph build
 * Replaces `adj.nonzero().t()` (layers.py:129, rule 0: entry != 0) and the `adj > 0`
 * mask (layers.py:41, rule 1: entry > 0).  adj is a dense [n, n] fp32 matrix with
 * arbitrary element strides (utils.py:55 yields column-major, load_data_ppi.py:86
 * row-major).  Step 1 writes rowptr (inclusive scan of per-row counts, rowptr[0]=0);
 * the caller reads E = rowptr[n], allocates col, and runs step 2. */
GATK_API size_t gatk_scan_workspace_bytes(int64_t n);
GATK_API int gatk_csr_from_dense_rowptr(const float* adj, int64_t n, int64_t row_stride, int64_t col_stride,
                               int rule, int64_t* rowptr, void* ws, size_t ws_bytes, void* stream);
GATK_API int gatk_csr_from_dense_fill(const float* adj, int64_t n, int64_t row_stride, int64_t col_stride,
                             int rule, const int64_t* rowptr, int32_t* col, void* stream);
/* COO (row-major sorted, as adj.nonzero() returns it) -> rowptr; col is coo_col cast. */
GATK_API int gatk_csr_from_coo(const int64_t* coo_row, const int64_t* coo_col, int64_t e, int64_t n,
                      int64_t* rowptr, int32_t* col, void* ws, size_t ws_bytes, void* stream);
/* Transposed pattern (source-major) with the permutation back into CSR edge ids; stable,
 * so destinations appear ascending within a source.  n_rows/n_cols: CSR shape. */
GATK_API size_t gatk_transpose_workspace_bytes(int64_t n_rows, int64_t n_cols, int64_t e);
GATK_API int gatk_csr_transpose(int64_t n_rows, int64_t n_cols, int64_t e, const int64_t* rowptr,
                       const int32_t* col, int64_t* tptr, int32_t* trow, int32_t* perm,
                       void* ws, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------ dropout (F.dropout, layers.py:34,37,43 / :132,136,153)
 * keep[i] = 1 with probability 1-p (Philox4x32-10 keyed by seed, counter = offset + i/4).
 * The layer kernels take a dropout site EITHER as a materialised keep mask (uint8, how the parity tests inject the
 * masks the reference drew) OR as (seed, drop_offset, p_drop) with the mask pointer NULL: the kernel then evaluates
 * exactly this stream where it consumes the decision (element index = the mask's flat index), so training with
 * p > 0 never materialises a mask.  p_drop = 0 with a NULL mask: site not dropped. */
GATK_API int gatk_dropout_keep_mask(uint8_t* keep, int64_t n, float p, uint64_t seed, uint64_t offset, void* stream);
/* y = keep ? x * scale : 0  (rows x cols with leading dims; keep dense rows*cols). */
GATK_API int gatk_mask_scale(const float* x, int64_t ldx, const uint8_t* keep, float scale, float* y, int64_t ldy,
                    int64_t rows, int64_t cols, void* stream);

/* ------------------------------------------------------------------ K1/K5: projection GEMMs
 * Row-major fp32 C[M,N] = op(A)[M,K] * op(B)[K,N] (+ C if accumulate), fp32-accurate.
 * Replaces torch.mm(h, W) / mm(h, skip_projection) (layers.py:35,48,134,166) and their
 * autograd (dW = h^T dWh, dh = dWh W^T).  transA: A stored [K,M]; transB: B stored [N,K].
 * ws: split-K scratch (gatk_gemm_workspace_bytes). */
GATK_API size_t gatk_gemm_workspace_bytes(int transA, int transB, int64_t M, int64_t N, int64_t K);
/* 1 when gatk_gemm would take the tcgen05/TMA 3xTF32 kernel for 16-byte aligned A and C (large
 * NN products whose row pitches are multiples of 16 bytes), 0 when it takes the fp32 SIMT kernel. */
GATK_API int gatk_gemm_uses_tensor_cores(int transA, int transB, int64_t M, int64_t N, int64_t K, int64_t lda,
                                         int64_t ldc, int accumulate);
GATK_API int gatk_gemm(int transA, int transB, int64_t M, int64_t N, int64_t K, const float* A, int64_t lda,
              const float* B, int64_t ldb, float* C, int64_t ldc, int accumulate,
              void* ws, size_t ws_bytes, void* stream);

/* Batched (per-head) products: for b in [0, batches)  C_b = op(A_b) op(B_b)  with A_b = A + b*a_bs, B_b = B + b*b_bs,
 * C_b = C + b*c_bs (element offsets: the batches are column blocks of the same matrices, or separate matrices).
 * epilogue 1 applies ELU to C (F.elu, layers.py:51,170).  Never accumulates.  One launch of the tcgen05 kernels
 * (A operand split into tensor memory, fp32 parity through the 3xTF32 scheme) covers all batches when
 *   !transA:           M >= 1024, K <= 512 and (N <= 64 or K <= 128), pitches / batch strides multiples of 4 floats;
 *   transA, !transB:   K >= 2048, 8 <= M, N <= 512 (deterministic split-K over all SMs);
 * other shapes run one gatk_gemm per batch.
 * elu_out (optional): the backward products of a layer with an ELU take dh' = gout * ELU'(v) as an operand (A of the
 * NT product dxagg_h = dh'_h W_h^T, B of the TN product dW_h = xagg_h^T dh'_h).  With elu_out = ELU(v) (the layer's
 * activated output, same shape and batch stride as that operand, row pitch ld_elu) the kernels load both tiles and
 * form the product while they split the operand, so dh' is never written to memory (F.elu's autograd, layers.py:51,170).
 * Only on the tensor-core paths: gatk_gemm_batched_fuses_elu_grad(...) == 1 for the shape; an error otherwise. */
GATK_API size_t gatk_gemm_batched_workspace_bytes(int transA, int transB, int64_t M, int64_t N, int64_t K, int batches);
GATK_API int gatk_gemm_batched_fuses_elu_grad(int transA, int transB, int64_t M, int64_t N, int64_t K, int batches,
                                              int64_t lda, int64_t a_bs, int64_t ldb, int64_t b_bs, int64_t ldc, int64_t c_bs);
GATK_API int gatk_gemm_batched(int transA, int transB, int64_t M, int64_t N, int64_t K, int batches, const float* A,
                               int64_t lda, int64_t a_bs, const float* B, int64_t ldb, int64_t b_bs, float* C, int64_t ldc,
                               int64_t c_bs, int epilogue, const float* elu_out, int64_t ld_elu, void* ws, size_t ws_bytes,
                               void* stream);

/* Projection with the reference's PER-HEAD input dropout (every head module drops the layer input with its own mask,
 * layers.py:34,132, and its skip projection uses the same dropped input, :48,166), all heads in ONE launch and
 * without materialising masks: the keep decision of element (h, i, k) of the [H, n, F] site is evaluated from
 * (seed, drop_offset) where x[i, k] is loaded.  W is [F, ldw] with the weight block of head h at columns
 * [h*Dp, (h+1)*Dp) and, when has_skip, the skip block at [H*Dp + h*Dp, ...); z / dz use the same column layout.
 *   fwd   z[i, blk_h] = sum_k x[i,k] m_h[i,k] inv_keep W[k, blk_h]
 *   dW    dW[k, blk_h] = sum_i x[i,k] m_h[i,k] inv_keep dz[i, blk_h]           (deterministic two-stage reduction;
 *         ws floats: gatk_gemm_heads_dropout_ws_floats)
 *   dx    dx[i,k] = inv_keep sum_h m_h[i,k] sum_{c in blk_h} dz[i,c] W[k,c]
 * fp32 SIMT kernels (exact products): the shapes that train with dropout are the citation graphs (n <= 2e4). */
GATK_API size_t gatk_gemm_heads_dropout_ws_floats(int64_t n, int F, int H, int Dp, int has_skip);
GATK_API int gatk_gemm_heads_dropout_fwd(int64_t n, int F, int H, int Dp, int has_skip, const float* x, int64_t ldx,
                                         const float* W, int64_t ldw, float* z, int64_t ldz, uint64_t seed,
                                         uint64_t drop_offset, float p_drop, void* stream);
GATK_API int gatk_gemm_heads_dropout_dw(int64_t n, int F, int H, int Dp, int has_skip, const float* x, int64_t ldx,
                                        const float* dz, int64_t ldz, float* dW, int64_t lddw, float* ws,
                                        uint64_t seed, uint64_t drop_offset, float p_drop, void* stream);
GATK_API int gatk_gemm_heads_dropout_dx(int64_t n, int F, int H, int Dp, int has_skip, const float* dz, int64_t ldz,
                                        const float* W, int64_t ldw, float* dx, int64_t lddx, uint64_t seed,
                                        uint64_t drop_offset, float p_drop, void* stream);

/* Attention-logit halves f_i = Wh_i . a[:D], g_j = Wh_j . a[D:] (layers.py:60-61, :141-144),
 * after the post-projection dropout (layers.py:37,136) applied IN PLACE to wh when
 * keep_wh != NULL (keep_wh is [n, H*Dp] dense).  a_src / a_dst are [H, Dp]. */
GATK_API int gatk_logits_fwd(int64_t n, int H, int Dp, float* wh, int64_t ldw, const uint8_t* keep_wh,
                    float inv_keep, const float* a_src, const float* a_dst, float* f, float* g,
                    uint64_t seed, uint64_t drop_offset, float p_drop, void* stream);

/* ------------------------------------------------------------------ K2: fused attention forward
 * One pass per destination row over its CSR edges: LeakyReLU(f_i + g_j), online
 * max-subtracted softmax, attention dropout, weighted aggregation of Wh_j, division by the
 * row sum, optional skip add and ELU.  Replaces layers.py:40-51 and :141-170 including
 * torch_scatter.scatter_max (:145) and both SpecialSpmm calls (:150,:156).
 *
 * Rows are local [0, n_dst); col indexes wh / g (global sources).  Rows longer than
 * seg_len are listed in hub_rows and processed as segments (hub_seg_ptr: int32
 * [n_hub+1], exclusive scan of per-hub segment counts) with partial softmax states in
 * hub_scratch (gatk_hub_scratch_floats) merged by a second kernel.
 * keep_att: [E, H] or NULL.  hagg (pre-skip, pre-ELU aggregation) and lse
 * (m + log l per row/head) are saved for backward; either may be NULL.
 * counter: one int32 of scratch (dynamic scheduler).  item_ptr (int32 [n_items+1], optional): edge-balanced
 * work items, item k = rows [item_ptr[k], item_ptr[k+1]); a warp claims one item per atomic, so skewed
 * degree distributions do not leave a few warps with all the long rows.  NULL: 8 rows per claim.
 * gatk_hub_scratch_floats(which, ...): floats of hub_scratch for which = 0 (attn_fwd),
 * 1 (attn_bwd_fused), 2 (attn_bwd_finish). */
GATK_API size_t gatk_hub_scratch_floats(int which, int H, int Dp, int n_hub_seg);
GATK_API int gatk_attn_fwd(int64_t n_dst, const int64_t* rowptr, const int32_t* col, int H, int Dp,
                  const float* wh, int64_t ldw, const float* f, int64_t ldf, const float* g, int64_t ldg,
                  const uint8_t* keep_att, float inv_keep, float alpha,
                  const float* skipv, int64_t lds, int act_elu,
                  float* hagg, float* out, int64_t ldo, float* lse,
                  int seg_len, const int32_t* hub_rows, const int32_t* hub_seg_ptr,
                  int n_hub, int n_hub_seg, float* hub_scratch, int32_t* counter,
                  const int32_t* item_ptr, int n_items, uint64_t seed, uint64_t drop_offset, float p_drop,
                  void* stream);

/* ------------------------------------------------------------------ K3/K4: backward of the fused attention
 * Autograd of layers.py:141-160, with the reference's dense N x N SpecialSpmmFunction.backward
 * (layers.py:81-90) replaced by O(E*D) CSR work that gathers each feature row ONCE:
 *
 *  prep    per destination row i, one contiguous RECORD of ldrec = gatk_attn_bwd_record_ld(H, Dp)
 *          floats:  [ dhp_i = dL/dh'_i = gout_i * ELU'(out_i)  |  (f_i, lse_i, c_i, 0) per head ]
 *          with c[i,h] = dhp_i . hagg_i (the softmax-backward row term).  out (the activated output)
 *          is needed only when a skip term was added before the ELU; with out == NULL the
 *          derivative is taken from hagg itself.  dhp2 (optional) receives a second copy of dhp
 *          (it is also dL/d(skip)).
 *  fused   per SOURCE row j over the transposed pattern (scatter-free): gathers record i once
 *          per edge and uses it for both  dz_ij = alpha_ij (keep/(1-p) dhp_i.Wh_j - c_i)
 *          LeakyReLU'(f_i+g_j)  and  dwh_j = sum_i alpha~_ij dhp_i + dg_j a_dst,  dg_j = sum_i
 *          dz_ij.  dz is written in CSR edge order (edge_dz [E,H]); keep_att is in CSR edge order.
 *          With df_acc (float [n_dst, lddf_acc], ZERO-INITIALISED) the kernel adds dz_ij straight into df_acc[i, h]
 *          (red.global.add) instead: edge_dz, perm and gatk_attn_bwd_finish are then not needed.  That is the form a
 *          SOURCE-row shard uses (tptr / trow cover this rank's sources, rec holds every destination: the entries of
 *          a destination row are spread over the ranks, their df partials are reduce-scattered by the caller).
 *          hub_* describe the TRANSPOSED pattern's long rows (scratch: which = 1).
 *  finish  per destination row: df_i = sum_j dz_ij (segmented sum over CSR rows), then
 *          dwh_i += df_i a_src and the post-projection dropout mask keep_wh (layers.py:37,136).
 *          hub_* describe the CSR pattern's long rows (scratch: which = 2).
 * f, g, df, dg are [rows, H] with explicit row pitches (ldf*, ldg, lddf, lddg): they may be column blocks of
 * the projection output / its gradient.  a_dst == NULL (fused) and a_src == NULL (finish) select the FOLDED
 * form used when no dropout sits between projection and logits: f = x (W a_src), g = x (W a_dst) come out of
 * the projection GEMM as extra columns, df / dg go back into its backward as extra columns, and the
 * dg a_dst / df a_src terms and the da reduction become parameter-sized algebra on the host side. */
GATK_API int64_t gatk_attn_bwd_record_ld(int H, int Dp);
GATK_API int gatk_attn_bwd_prep(int64_t n, int H, int Dp, const float* gout, int64_t ldgo, const float* out,
                                int64_t ldo, int act_elu, const float* hagg, int64_t ldh, const float* f,
                                int64_t ldf, const float* lse, float* rec, int64_t ldrec, float* dhp2, int64_t lddhp2,
                                void* stream);
GATK_API int gatk_attn_bwd_fused(int64_t n_src, const int64_t* tptr, const int32_t* trow, const int32_t* perm,
                                 int H, int Dp, const float* wh, int64_t ldw, const float* g, int64_t ldg,
                                 const float* rec, int64_t ldrec, const uint8_t* keep_att, float inv_keep,
                                 float alpha, const float* a_dst, float* dwh, int64_t lddwh, float* dg,
                                 int64_t lddg, float* edge_dz, float* df_acc, int64_t lddf_acc,
                                 int seg_len, const int32_t* hub_rows, const int32_t* hub_seg_ptr, int n_hub,
                                 int n_hub_seg, float* hub_scratch, int32_t* counter,
                                 const int32_t* item_ptr, int n_items, uint64_t seed, uint64_t drop_offset,
                                 float p_drop, void* stream);
GATK_API int gatk_attn_bwd_finish(int64_t n, const int64_t* rowptr, int H, int Dp, const float* edge_dz,
                                  const float* a_src, const uint8_t* keep_wh, float inv_keep, float* dwh,
                                  int64_t lddwh, float* df, int64_t lddf, int seg_len, const int32_t* hub_rows,
                                  const int32_t* hub_seg_ptr, int n_hub, int n_hub_seg, float* hub_scratch,
                                  uint64_t seed, uint64_t drop_offset, float p_drop, void* stream);

/* da_src[h,:] = sum_i df[i,h] Wh[i,h,:],  da_dst[h,:] = sum_j dg[j,h] Wh[j,h,:]  (autograd of
 * layers.py:60-61 / :144).  ws floats: gatk_da_workspace_floats(H, Dp). Deterministic. */
GATK_API size_t gatk_da_workspace_floats(int H, int Dp);
GATK_API int gatk_da_reduce(int64_t n, int H, int Dp, const float* wh, int64_t ldw, const float* df,
                   const float* dg, float* da_src, float* da_dst, float* ws, void* stream);

/* ------------------------------------------------------------------ aggregate-first form (narrow inputs)
 * The same layer as K1-K5 (layers.py:134-170, models.py:29-35) with the neighbour sum taken BEFORE the
 * projection:  h'_ih = (sum_j alpha_ijh x_j) W_h.  Valid when no dropout sits between the projection and
 * the logits; pays when the input row is narrower than the projected row (F_in < H*D: the products shape
 * gathers 400 B instead of 2 KiB per stored entry).
 *
 *  logits_pack  f = x u, g = x v with [u | v] = [W a_src | W a_dst] (uv is [F, >= 2H], u in columns 0..H-1,
 *          v in H..2H-1) and the GATHER ROWS  xg_i = [x_i (F floats, zero padded to Fp = 4*ceil(F/4)) |
 *          g_i (H floats) | zero pad]  with row pitch >= gatk_xg_pitch(Fp, H) = Fp + 4*ceil(H/4) floats (rows
 *          16-byte aligned for the bulk copies; a stored entry costs one gather for both x_j and g_j).
 *  x_fwd   per destination row: softmax_j LeakyReLU(f_i + g_j) (online, max-subtracted) and
 *          xagg[i, h*Fp:(h+1)*Fp] = sum_j alpha_ijh x_j;  lse[i,h] = m + log l.  The caller then projects:
 *          out_h = xagg_h W_h (+ x S_h) through gatk_gemm and applies gatk_elu_fwd.
 *  x_bwd   per destination row, given dxagg_ih = dh'_ih W_h^T (gatk_gemm):  c_ih = dxagg_ih . xagg_ih,
 *          ds_ijh = alpha_ijh (dxagg_ih . x_j - c_ih) LeakyReLU'(f_ih + g_jh)  ->  ds [E, H] in CSR edge order,
 *          df[i,h] = sum_j ds_ijh.  With iperm (int32 [E], CSR entry -> position in the transposed pattern, the
 *          inverse of gatk_csr_transpose's perm) ds is written in TRANSPOSED order, so that edge_tsum streams
 *          it (perm = NULL there) instead of gathering 32-byte pieces.  (dW_h = xagg_h^T dh'_h is a gatk_gemm;
 *          dx is not produced: the form is for layers whose input needs no gradient.)
 *          With dg_acc (float [n_src, lddg_acc], ZERO-INITIALISED by the caller) the kernel adds every ds_ijh
 *          straight into dg_acc[j, h] with red.global.add (vector reductions into an L2-resident array): ds may
 *          then be NULL and gatk_edge_tsum is not needed.  The sum order is then scheduling dependent (fp32
 *          rounding differences of ~1e-7 in dg); ds + gatk_edge_tsum is the bit-reproducible route.
 *  edge_tsum  dg[j,h] = sum_i ds_ijh: segmented sum along the transposed pattern (tptr, perm from
 *          gatk_csr_transpose; perm = NULL when ds is already in transposed order); sources with more than long_len entries are listed in long_rows.
 * n_src: rows of xg (sources; col indexes them), n_dst: destination rows of this shard.
 * Hub rows (longer than seg_len) as in gatk_attn_fwd; scratch floats: gatk_attn_x_scratch_floats(which, ...)
 * with which = 0 (x_fwd), 1 (x_bwd).  H <= 8, Fp <= 512, H_pow2 * ceil((Fp/4 + ceil(H/4)) / 32) <= 16. */
GATK_API int64_t gatk_xg_pitch(int Fp, int H);
GATK_API int gatk_logits_pack(int64_t n, int F, int H, const float* x, int64_t ldx, const float* uv, int64_t lduv,
                              float* xg, int64_t ldxg, float* f, int64_t ldf, void* stream);
/* The same kernel as the forward exchange of a destination-row shard group: besides xg (this GPU's own copy of
 * its rows) every packed row is also written to peer_xg[q] + row*ldxg, q < n_peers <= GATK_MAX_PEERS, pointers
 * into the other GPUs' gathered-row buffers mapped into this process (CUDA peer / symmetric memory), at the row
 * that corresponds to local row 0.  whole_rows = 0 pushes only the g columns (the input columns of the peers'
 * copies are already in place: a static first-layer input), 1 the whole row.  The caller orders the exchange
 * (a barrier over the group before the rows are overwritten and after this call).  There is no reference
 * counterpart: the reference is single-device. */
#define GATK_MAX_PEERS 15
GATK_API int gatk_logits_pack_push(int64_t n, int F, int H, const float* x, int64_t ldx, const float* uv, int64_t lduv,
                                   float* xg, int64_t ldxg, float* f, int64_t ldf, int n_peers, float* const* peer_xg,
                                   int whole_rows, void* stream);
GATK_API size_t gatk_attn_x_scratch_floats(int which, int H, int Fp, int n_hub_seg);
GATK_API int gatk_attn_x_fwd(int64_t n_src, int64_t n_dst, const int64_t* rowptr, const int32_t* col, int H, int Fp,
                             const float* xg, int64_t ldxg, const float* f, int64_t ldf, float alpha,
                             float* xagg, int64_t ldxa, float* lse, int seg_len, const int32_t* hub_rows,
                             const int32_t* hub_seg_ptr, int n_hub, int n_hub_seg, float* hub_scratch,
                             int32_t* counter, const int32_t* item_ptr, int n_items, void* stream);
GATK_API int gatk_attn_x_bwd(int64_t n_src, int64_t n_dst, const int64_t* rowptr, const int32_t* col, int H, int Fp,
                             const float* xg, int64_t ldxg, const float* f, int64_t ldf, const float* lse, float alpha,
                             const float* xagg, int64_t ldxa, const float* dxagg, int64_t ldd, float* ds,
                             const int32_t* iperm, float* dg_acc, int64_t lddg_acc, float* df, int64_t lddf, int seg_len,
                             const int32_t* hub_rows, const int32_t* hub_seg_ptr, int n_hub,
                             int n_hub_seg, float* hub_scratch, int32_t* counter, const int32_t* item_ptr,
                             int n_items, void* stream);
GATK_API int gatk_edge_tsum(int64_t n_src, const int64_t* tptr, const int32_t* perm, int H, const float* ds,
                            float* dg, int64_t lddg, int long_len, const int32_t* long_rows, int n_long,
                            void* stream);
/* ELU in place on [n, cols] (F.elu, layers.py:51,170) and dhp = gout * ELU'(v) taken from out = ELU(v). */
GATK_API int gatk_elu_fwd(int64_t n, int64_t cols, float* buf, int64_t ld, void* stream);
GATK_API int gatk_elu_bwd(int64_t n, int64_t cols, const float* gout, int64_t ldg, const float* out, int64_t ldo,
                          float* dhp, int64_t ldd, void* stream);

/* ------------------------------------------------------------------ GATv2 flavour: SpGraphAttentionLayerV2 (layers.py:255-313)
 * z [n_src, ldz] = [Whi | Whj | (skip)] (H*Dp columns each: the two projections of the input, layers.py:265-266, and the
 * optional skip projection :301), a [H, Dp].  Per stored entry (i, j):  s_ij = a . LeakyReLU(Whi_i + Whj_j)
 * (layers.py:275-278), softmax over the row (scatter_max / exp / rowsum, :280-288), attention dropout through
 * keep_att [E, H] (:289), h'_i = sum_j alpha~_ij Whi_j (:291-295), + skip, ELU (:301-305).  One fused CSR pass per
 * destination row; hagg (pre-skip, pre-ELU) and lse are saved for backward.
 * Backward (one destination-major pass): dz [n_src, ldz] and da [H, Dp] must be ZERO-INITIALISED; source-side
 * gradients are accumulated with vector reductions (fp32 addition order is scheduling dependent, ~1e-7). */
GATK_API int gatk_attn_v2_fwd(int64_t n_dst, const int64_t* rowptr, const int32_t* col, int H, int Dp, const float* z,
                              int64_t ldz, const float* a, const uint8_t* keep_att, float inv_keep, float alpha,
                              int has_skip, int act_elu, float* hagg, float* out, int64_t ldo, float* lse,
                              int32_t* counter, void* stream);
GATK_API int gatk_attn_v2_bwd(int64_t n_dst, const int64_t* rowptr, const int32_t* col, int H, int Dp, const float* z,
                              int64_t ldz, const float* a, const uint8_t* keep_att, float inv_keep, float alpha,
                              int has_skip, int act_elu, const float* hagg, const float* out, int64_t ldo,
                              const float* lse, const float* gout, int64_t ldgo, float* dz, float* da, int32_t* counter,
                              void* stream);

/* ------------------------------------------------------------------ K6: head combine (models.py:32-34)
 * mode 0: strip the Dp padding -> out [n, H*D] (torch.cat);  mode 1: mean over heads ->
 * out [n, D] (torch.mean(torch.stack)).  The backward scatters gout back to [n, H*Dp]. */
GATK_API int gatk_head_combine(int64_t n, int H, int D, int Dp, const float* in, int64_t ldi, int mode,
                      float* out, void* stream);
GATK_API int gatk_head_combine_bwd(int64_t n, int H, int D, int Dp, const float* gout, int mode,
                          float* gin, int64_t ldi, void* stream);

/* ------------------------------------------------------------------ loss heads + metrics of the callers' step
 * Citation scripts (train.py:151-160): loss = nll_loss(log_softmax(elu(logits))[idx], labels[idx]), accuracy of the
 * same rows (utils.py:92-96).  fwd adds into stats (double[2], zeroed by the caller): [0] the SUM of the per-row
 * losses, [1] the number of correct predictions.  bwd ADDS scale * gscale[0] * dloss/dlogits for the selected rows into
 * dlogits (zero-initialised; gscale: optional device scalar, the upstream gradient of the mean loss).
 * idx: int64 [n_idx] row ids or NULL for rows 0..n_idx-1; labels: int64 [N].
 * PPI script (train_ppi.py:106-120): BCEWithLogits + micro-F1 of (logits > 0): stats (double[4]) += [sum of the
 * element losses, TP, FP, FN]; micro-F1 = 2TP / (2TP + FP + FN).  Nothing here synchronises or leaves the device. */
GATK_API int gatk_nll_head_fwd(int64_t n_idx, const int64_t* idx, const float* logits, int64_t ld, const int64_t* labels,
                               int C, double* stats, void* stream);
GATK_API int gatk_nll_head_bwd(int64_t n_idx, const int64_t* idx, const float* logits, int64_t ld, const int64_t* labels,
                               int C, const float* gscale, float scale, float* dlogits, int64_t ldd, void* stream);
GATK_API int gatk_bce_f1_fwd(int64_t total, const float* logits, const float* labels, double* stats, void* stream);
GATK_API int gatk_bce_bwd(int64_t total, const float* logits, const float* labels, const float* gscale, float scale,
                          float* dlogits, void* stream);

/* ------------------------------------------------------------------ SpecialSpmm (layers.py:70-95)
 * out[n_rows, k] = COO(row, col, val) @ b ; grad_val[e] = gout[row_e] . b[col_e] ;
 * grad_b = COO^T @ gout.  Indices int64 [E] each, any order.  out / grad_b must be
 * zero-filled by the caller (scatter with float atomics). */
GATK_API int gatk_spmm_coo_fwd(const int64_t* row, const int64_t* col, const float* val, int64_t e, int64_t k,
                      const float* b, float* out, void* stream);
GATK_API int gatk_spmm_coo_bwd(const int64_t* row, const int64_t* col, const float* val, int64_t e, int64_t k,
                      const float* b, const float* gout, float* grad_val, float* grad_b, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GATK_H */
