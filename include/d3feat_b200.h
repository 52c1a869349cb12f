/*
 * d3feat_b200.h -- C ABI of libd3feat_b200.so: the B200 (sm_100a) implementation of
 * D3Feat's data-parallel hot path.  Plain pointers and sizes only; no torch types.
 *
 * Conventions:
 *   - every array pointer is a DEVICE pointer unless the name ends in _host;
 *   - the caller owns every buffer (outputs and workspace); no entry point allocates
 *     device memory or synchronises the device; all work is enqueued on `stream`
 *     (a cudaStream_t passed as void*);
 *   - all matrices are dense row-major; indices are int32 or int64 as flagged;
 *   - return value: D3F_OK (0) or a negative d3f_status; d3f_last_error_string()
 *     describes the last failure on the calling thread;
 *   - thread-safe for calls that use distinct streams and distinct buffers: nothing in this header reads or writes
 *     process-global state (the A/B selectors and measurement hooks of the test suite live in d3feat_b200_debug.h
 *     and are not part of the product ABI).
 *
 * Each entry point cites the reference interface it replaces (file:line under the
 * XuyangBai/D3Feat.pytorch tree).  INTEGRATION.md shows the reference-side binding.
 */
#ifndef D3FEAT_B200_H
#define D3FEAT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* d3f_stream; /* cudaStream_t */

typedef enum {
    D3F_OK = 0,
    D3F_ERR_INVALID = -1,     /* bad argument (null pointer, negative size, unknown mode) */
    D3F_ERR_CUDA = -2,        /* a CUDA runtime call failed; see d3f_last_error_string() */
    D3F_ERR_WORKSPACE = -3,   /* workspace smaller than *_workspace_bytes() */
    D3F_ERR_UNSUPPORTED = -4  /* shape outside what the kernels implement */
} d3f_status;

/* influence / aggregation / metric / loss selectors (values of the reference's config strings) */
enum { D3F_INFLUENCE_CONSTANT = 0, D3F_INFLUENCE_LINEAR = 1, D3F_INFLUENCE_GAUSSIAN = 2 }; /* blocks.py:329-343 */
enum { D3F_AGGREGATION_SUM = 0, D3F_AGGREGATION_CLOSEST = 1 };                             /* blocks.py:346-351 */
enum { D3F_METRIC_EUCLIDEAN = 0, D3F_METRIC_SQEUCLIDEAN = 1, D3F_METRIC_CITYBLOCK = 2,
       D3F_METRIC_COSINE = 3, D3F_METRIC_ARCCOSINE = 4 };                                  /* loss.py:8-44 */
enum { D3F_LOSS_CIRCLE = 0, D3F_LOSS_CONTRASTIVE = 1 };                                    /* loss.py:47,100 */

int d3f_version(void);
const char* d3f_last_error_string(void);
/* number of CUDA kernels this library has launched in this process (diagnostic counter) */
unsigned long long d3f_launch_count(void);

/* ------------------------------------------------------------------------------------------
 * Radius neighbours.  Replaces radius_neighbors.batch_query
 * (cpp_wrappers/cpp_neighbors/wrapper.cpp:58-238 -> neighbors/neighbors.cpp:211-332), called
 * from datasets/dataloader.py:63.
 *
 * For every query, all supports OF THE SAME BATCH ELEMENT with fp32 d2 < radius*radius
 * (d2 = (dx*dx + dy*dy) + dz*dz, no FMA), ordered by ascending (d2, index); indices are global
 * rows of `supports`; rows are padded with n_supports.  Only the first `max_cols` neighbours of
 * a row are written (dataloader.py:64-65 truncation).
 *
 *   queries   [n_queries,3] f32     supports [n_supports,3] f32
 *   q_lengths [n_batch] i32         s_lengths [n_batch] i32      (device)
 *   n_queries / n_supports are the ALLOCATED row counts; the real rows are the first sum(q_lengths) /
 *   sum(s_lengths) ones.  They may be smaller (static capacities for CUDA-graph capture): padding query
 *   rows get an all-shadow row, padding support rows are never returned.
 *   pad_index: the shadow index written into unused slots (-1: n_supports, the reference's value)
 *   out_idx   [n_queries,max_cols] i32 (idx_is_64=0) or i64 (idx_is_64=1); may be NULL: count only
 *   out_info  [4] i32 (device): [0] = max neighbour count over all queries BEFORE truncation
 *             (the reference's matrix width, neighbors.cpp:300), [1] = 1 if some row had more
 *             than `row_capacity` in-range supports (its selection is then incomplete: retry with a
 *             larger capacity), [2] = number of occupied grid cells, [3] reserved.
 *   row_capacity: per-query candidate buffer (power of two, 64..8192).
 */
size_t d3f_radius_neighbors_workspace_bytes(int n_queries, int n_supports, int n_batch);
int d3f_radius_neighbors(const float* queries, const float* supports,
                         const int32_t* q_lengths, const int32_t* s_lengths, int n_batch,
                         int n_queries, int n_supports, float radius, int max_cols,
                         void* out_idx, int idx_is_64, int pad_index, int32_t* out_info, int row_capacity,
                         void* workspace, size_t workspace_bytes, d3f_stream stream);

/* ------------------------------------------------------------------------------------------
 * Grid subsampling (barycentre per occupied voxel, points only).  Replaces
 * grid_subsampling.subsample_batch (cpp_wrappers/cpp_subsampling/wrapper.cpp:62-333 ->
 * grid_subsampling/grid_subsampling.cpp:109-211), called from datasets/dataloader.py:17.
 * Output values AND order are bit-identical to the reference, i.e. the iteration order of
 * libstdc++'s std::unordered_map<size_t,SampledData> (grid_subsampling.cpp:48,85).
 *
 *   points [n_points,3] f32 (n_points = allocated rows, real rows = sum(lengths)), lengths [n_batch] i32 (device)
 *   out_points  [out_capacity,3] f32; the first sum(out_lengths[0..n_batch)) rows are valid, the rest is zeroed
 *   out_lengths [n_batch + 1] i32 (device): per-cloud counts (-1: voxel grid beyond the supported key range);
 *               [n_batch] = 1 if out_capacity was too small (the output is then truncated)
 */
size_t d3f_grid_subsample_workspace_bytes(int n_points, int n_batch);
int d3f_grid_subsample(const float* points, const int32_t* lengths, int n_batch, int n_points,
                       float sample_dl, float* out_points, int out_capacity, int32_t* out_lengths,
                       void* workspace, size_t workspace_bytes, d3f_stream stream);

/* ------------------------------------------------------------------------------------------
 * KPConv.  Replaces the ATen op chain of KPConv.forward (models/blocks.py:237-382) and its
 * autograd backward.
 *
 *   q_pts [Nq,3], s_pts [Ns,3], inds [Nq,H] (i32 or i64, row stride ld_inds elements, value Ns =
 *   shadow neighbour), x [Ns,Cin], weights [K,Cin,Cout], kernel_points [K,3] (rigid) or
 *   [Nq,K,3] (deformed != 0: blocks.py:286-291, which also enables the in-range neighbour filter
 *   of blocks.py:300-324), modulations [Nq,K] or NULL (blocks.py:365-366).
 *
 * forward outputs:
 *   out [Nq,Cout]; wf [Nq,K,Cin] = kernel-point-weighted neighbour features after modulation
 *   (the [n_points,n_kpoints,in_fdim] tensor of blocks.py:362-366; saved for backward).  wf may be NULL for rigid,
 *   unmodulated, linear-influence, sum-aggregation layers with Cin = Cout = 32, K <= 15 and <= 48 neighbour columns:
 *   those run as ONE fused kernel (gather + correlation + tcgen05 contraction, csrc/kpconv_fused.cu) that never
 *   materialises wf; their backward then uses the transposed lists (d3f_kpconv_backward_ex);
 *   wf_unmod [Nq,K,Cin] or NULL (required iff modulations != NULL: wf before modulation);
 *   inv_n [Nq] = 1 / max(1, #neighbours with positive feature sum) (blocks.py:377-380);
 *   min_d2 [Nq,K] or NULL (deformed only, blocks.py:303).
 */
size_t d3f_kpconv_workspace_bytes(int n_queries, int n_supports, int n_neighbors, int K, int c_in, int c_out);
int d3f_kpconv_forward(const float* q_pts, const float* s_pts, const void* inds, int idx_is_64,
                       int64_t ld_inds, const float* x, const float* weights,
                       const float* kernel_points, int deformed, const float* modulations,
                       int n_queries, int n_supports, int n_neighbors, int K, int c_in, int c_out,
                       float kp_extent, int influence, int aggregation,
                       float* out, float* wf, float* wf_unmod, float* inv_n, float* min_d2,
                       void* workspace, size_t workspace_bytes, d3f_stream stream);

/* Same as d3f_kpconv_forward with the block epilogue fused into the contraction:
 *   out = act(conv + bias[Cout]),  act = LeakyReLU(slope) if leaky_relu != 0
 * (SimpleBlock / ResnetBottleneckBlock apply `x + bias` (BatchNormBlock with use_bn = False) and LeakyReLU(0.1) right
 * after the convolution, blocks.py:597, :671; the deformable offset head adds offset_bias, blocks.py:246). */
int d3f_kpconv_forward_ex(const float* q_pts, const float* s_pts, const void* inds, int idx_is_64,
                          int64_t ld_inds, const float* x, const float* weights,
                          const float* kernel_points, int deformed, const float* modulations,
                          int n_queries, int n_supports, int n_neighbors, int K, int c_in, int c_out,
                          float kp_extent, int influence, int aggregation,
                          const float* bias, int leaky_relu, float slope,
                          float* out, float* wf, float* wf_unmod, float* inv_n, float* min_d2,
                          void* workspace, size_t workspace_bytes, d3f_stream stream);

/* backward: grad_out [Nq,Cout] ->
 *   grad_x [Ns,Cin] or NULL (fully overwritten), grad_weights [K,Cin,Cout] or NULL (overwritten),
 *   grad_kernel_points [Nq,K,3] or NULL (deformed only), grad_modulations [Nq,K] or NULL. */
int d3f_kpconv_backward(const float* q_pts, const float* s_pts, const void* inds, int idx_is_64,
                        int64_t ld_inds, const float* x, const float* weights,
                        const float* kernel_points, int deformed, const float* modulations,
                        int n_queries, int n_supports, int n_neighbors, int K, int c_in, int c_out,
                        float kp_extent, int influence, int aggregation,
                        const float* wf, const float* wf_unmod, const float* inv_n,
                        const float* grad_out,
                        float* grad_x, float* grad_weights, float* grad_kernel_points,
                        float* grad_modulations,
                        void* workspace, size_t workspace_bytes, d3f_stream stream);

/* Transposed neighbour lists for the atomic-free backward: for every support j the queries i with inds[i, h] = j
 * (any h), as CSR:  t_offsets [n_supports + 1] i32,  t_src [n_queries * n_neighbors capacity] i32 (query indices in
 * ascending order inside a list, so the backward that sums in list order is bit-reproducible).  One call per neighbour matrix; every KPConv layer that uses the matrix shares it.
 * d3f_kpconv_backward_ex = d3f_kpconv_backward plus (t_offsets, t_src): when both are given and the layer is unmodulated
 * with Cout % 32 == 0, grad_x is computed as a forward-style gather over the lists followed by one GEMM with W^T -- no float
 * atomics (the 45 M reductions of level 0 cost 200 us however they are issued) and a deterministic result per list order.
 * Deformable layers (kernel_points [Nq,K,3]) take the same path with each listing query's own kernel points; their
 * grad_kernel_points (query-major) still comes from the scatter kernel, which then issues no grad_x reductions.
 * Otherwise it falls back to the scatter of d3f_kpconv_backward.  The KPConv workspace covers both. */
size_t d3f_neighbors_transpose_workspace_bytes(int n_supports);
int d3f_neighbors_transpose(const void* inds, int idx_is_64, int64_t ld_inds, int n_queries, int n_supports,
                            int n_neighbors, int32_t* t_offsets, int32_t* t_src, void* workspace, size_t workspace_bytes,
                            d3f_stream stream);
int d3f_kpconv_backward_ex(const float* q_pts, const float* s_pts, const void* inds, int idx_is_64,
                           int64_t ld_inds, const float* x, const float* weights,
                           const float* kernel_points, int deformed, const float* modulations,
                           int n_queries, int n_supports, int n_neighbors, int K, int c_in, int c_out,
                           float kp_extent, int influence, int aggregation,
                           const float* wf, const float* wf_unmod, const float* inv_n,
                           const float* grad_out,
                           float* grad_x, float* grad_weights, float* grad_kernel_points,
                           float* grad_modulations,
                           const int32_t* t_offsets, const int32_t* t_src,
                           void* workspace, size_t workspace_bytes, d3f_stream stream);

/* The list-based backward of a rigid layer in two steps (what d3f_kpconv_backward_ex does internally when the lists are
 * given), exposed so that the two GEMMs of step 2 can run concurrently on different streams:
 *   1. G [Ns, K, Cout] = sum over the queries i that list support j of  w(i,k,j) * inv_n[i] * grad_out[i, :]
 *      (the forward gather over the TRANSPOSED lists, reading rows of grad_out; Cout % 32 == 0);
 *   2. grad_x [Ns, Cin] = G x W^T  and / or  grad_weights [K, Cin, Cout] = x^T G   (either may be NULL).
 * Neither step needs the forward's kernel-point-weighted features wf: a training forward of such a layer may pass
 * wf = NULL to d3f_kpconv_forward_ex.  grad_weights_prezeroed != 0: the caller cleared grad_weights (see *_prezeroed). */
int d3f_kpconv_gather_transposed(const float* q_pts, const float* s_pts, const int32_t* t_offsets, const int32_t* t_src,
                                 const float* grad_out, const float* inv_n, const float* kernel_points,
                                 int n_queries, int n_supports, int K, int c_out, float kp_extent, int influence,
                                 int aggregation, float* G, d3f_stream stream);
int d3f_kpconv_grads_from_gathered(const float* G, const float* x, const float* weights, int n_supports, int K, int c_in,
                                   int c_out, float* grad_x, float* grad_weights, int grad_weights_prezeroed,
                                   d3f_stream stream);

/* ------------------------------------------------------------------------------------------
 * Pairwise descriptor distance + descriptor loss + detector loss.  Replaces cdist,
 * CircleLoss.forward / ContrastiveLoss.forward and DetLoss.forward (utils/loss.py:8-44, 111-141,
 * 55-97, 149-158) as wired by trainer.py:90-98.
 *
 *   anchor, positive [P,D] f32; dist_keypts [P,P] f64 (keypts_is_f64=1) or f32;
 *   anc_score, pos_score [P] f32 or NULL (then no detector loss);
 *   dists [P,P] out (the `dists` the reference returns: for contrastive it carries the +10 bumps);
 *   stats [8] out: [0] descriptor loss, [1] detector loss, [2] accuracy %, [3] mean furthest
 *   positive, [4] mean average negative; furthest_pos [P], avg_neg [P] out.
 *   aux: opaque buffer of d3f_pair_loss_aux_floats(P) floats, written by forward, read by backward.
 */
size_t d3f_pair_loss_aux_floats(int P);
int d3f_pair_dist(const float* a, const float* b, int Pa, int Pb, int D, int metric, float* dists,
                  d3f_stream stream);
int d3f_pair_loss_forward(const float* anchor, const float* positive, int P, int D,
                          const void* dist_keypts, int keypts_is_f64,
                          const float* anc_score, const float* pos_score,
                          int loss_kind, int metric, double safe_radius, float pos_margin,
                          float neg_margin, float log_scale,
                          float* dists, float* stats, float* furthest_pos, float* avg_neg,
                          float* aux, d3f_stream stream);
/* grad_losses [2] (device): dL/d(descriptor loss), dL/d(detector loss).
 * outputs: grad_anchor, grad_positive [P,D]; grad_anc_score, grad_pos_score [P] or NULL. */
int d3f_pair_loss_backward(const float* anchor, const float* positive, int P, int D,
                           const void* dist_keypts, int keypts_is_f64,
                           const float* anc_score, const float* pos_score,
                           int loss_kind, int metric, double safe_radius, float pos_margin,
                           float neg_margin, float log_scale,
                           const float* dists, const float* aux, const float* grad_losses,
                           float* grad_anchor, float* grad_positive,
                           float* grad_anc_score, float* grad_pos_score, d3f_stream stream);

/* Detector loss of utils/loss.py:149-158 on ANY [P,P] fp32 distance matrix (row stride ld) -- the drop-in form of
 * DetLoss.forward(dists, anc_score, pos_score) for matrices that do not come from d3f_pair_loss_forward.
 * loss [1] out; rowval [P] f32 and arg [2P] i32 are kept for the backward.  grad_loss [1] (device); grad_dists [P,P]
 * (zero-filled by the call) or NULL; grad_anc_score / grad_pos_score [P] or NULL. */
int d3f_det_loss_forward(const float* dists, int ld, const float* anc_score, const float* pos_score, int P,
                         float* loss, float* rowval, int32_t* arg, d3f_stream stream);
int d3f_det_loss_backward(const float* rowval, const int32_t* arg, const float* anc_score, const float* pos_score,
                          int P, const float* grad_loss, float* grad_dists, int ld, float* grad_anc_score,
                          float* grad_pos_score, d3f_stream stream);

/* ------------------------------------------------------------------------------------------
 * Multi-GPU exchange step (the reference has no multi-GPU code): pack a rank's P selected descriptor
 * pairs, scores and keypoint distances into ONE chunk for a single all-gather, and unpack W gathered chunks into the
 * inputs of the (W*P)^2 cross-fragment loss (dist_keypts block-diagonal, +inf between different fragments).
 * chunk = f32 anchor [P,D] | f32 positive [P,D] | f32 anc_score [P] | f32 pos_score [P] | f64 dist_keypts [P,P]. */
size_t d3f_exchange_chunk_bytes(int P, int D);
int d3f_exchange_pack(const float* anchor, const float* positive, const float* anc_score, const float* pos_score,
                      const void* dist_keypts, int dk_is_f64, int P, int D, void* chunk, d3f_stream stream);
int d3f_exchange_unpack(const void* all, int W, int P, int D, float* A, float* Pos, float* SA, float* SP,
                        double* DK, d3f_stream stream);

/* ------------------------------------------------------------------------------------------
 * Gathers between the KPConv layers (they dominate a GPU training step
 * when left to ATen advanced indexing).  Shadow index (>= n_supports) reads a zero row.
 *
 * max pool      : max_pool(x, inds) of models/blocks.py:94-110 (strided-block shortcut, MaxPoolBlock);
 *                 argmax [Nq,C] i32 = winning support row per output (-1 = shadow) for the backward.
 * gather rows   : closest_pool(x, inds) = x_pad[inds[:,0]] of blocks.py:79-91 (NearestUpsampleBlock) and the
 *                 correspondence row-selects of trainer.py:91-94; idx is read with an element stride.
 * detection     : KPFCNN.detection_scores (models/architectures.py:322-368), train (eval_mode=0) or test
 *                 (eval_mode=1: exact-equality local-max gate); features [N,C<=32], neighbors [N,H];
 *                 gmax_state: 16-byte device scratch written by forward and read by backward.
 * valid_width (device int32, may be NULL): the neighbour matrix may be allocated wider than the reference's
 *                 matrix (static capacities); columns >= *valid_width are treated as absent, not as shadow
 *                 neighbours (a row that fills the reference's matrix has no zero shadow row in its max).
 * Backward entry points fully overwrite their grad_* output.
 */
/* out[n] = sum over rows of x[n_rows, n_cols] (bias gradients of the fused UnaryBlock) */
int d3f_colsum(const float* x, int n_rows, int n_cols, float* out, d3f_stream stream);
int d3f_colsum_prezeroed(const float* x, int n_rows, int n_cols, float* out, d3f_stream stream);
/* LeakyReLU backward fused with the bias gradient: dz = grad * (y > 0 ? 1 : slope) with y the saved activation OUTPUT,
 * colsum[n] = sum_m dz[m, n].  n_cols must be 4 * 2^j with 16-byte aligned buffers, else D3F_ERR_UNSUPPORTED. */
int d3f_leaky_backward_colsum(const float* grad, const float* y, float slope, int n_rows, int n_cols, float* dz,
                              float* colsum, d3f_stream stream);
int d3f_leaky_backward_colsum_prezeroed(const float* grad, const float* y, float slope, int n_rows, int n_cols, float* dz,
                                        float* colsum, d3f_stream stream);
int d3f_max_pool_forward(const float* x, const void* inds, int idx_is_64, int64_t ld_inds, int n_queries,
                         int n_supports, int n_neighbors, int channels, const int32_t* valid_width, float* out,
                         int32_t* argmax, d3f_stream stream);
int d3f_max_pool_backward(const float* grad_out, const int32_t* argmax, int n_queries, int n_supports,
                          int channels, float* grad_x, d3f_stream stream);
int d3f_gather_rows_forward(const float* x, const void* idx, int idx_is_64, int64_t idx_stride, int n_rows,
                            int n_supports, int channels, float* out, d3f_stream stream);
int d3f_gather_rows_backward(const float* grad_out, const void* idx, int idx_is_64, int64_t idx_stride,
                             int n_rows, int n_supports, int channels, float* grad_x, d3f_stream stream);
int d3f_detection_scores_forward(const float* features, const void* neighbors, int idx_is_64, int64_t ld_inds,
                                 int n_points, int n_neighbors, int channels, int eval_mode,
                                 const int32_t* valid_width, float* scores, void* gmax_state, d3f_stream stream);
int d3f_detection_scores_backward(const float* features, const void* neighbors, int idx_is_64, int64_t ld_inds,
                                  int n_points, int n_neighbors, int channels, int eval_mode,
                                  const int32_t* valid_width, const void* gmax_state, const float* grad_scores,
                                  float* grad_features, d3f_stream stream);

/* ------------------------------------------------------------------------------------------
 * Mutual nearest-neighbour matching of two descriptor sets.  Replaces build_correspondence
 * (geometric_registration/common.py:5-21; numpy on the host in the reference, called from evaluate.py per fragment pair):
 *   distance = sqrt(2 - 2 * source @ target^T)   (fp32),
 *   source_arg[i] = argmin_j distance[i, j],  target_arg[j] = argmin_i distance[i, j]   (numpy.argmin semantics: the
 *   first NaN of a row / column wins -- 2 - 2<s,t> < 0 happens for descriptors a hair longer than 1 -- else the first
 *   minimum),  pairs = [(i, source_arg[i]) : target_arg[source_arg[i]] == i] in ascending i.
 *   source [n_source, dim], target [n_target, dim] f32;  source_arg [n_source], target_arg [n_target] i32 out;
 *   pairs [min(n_source, n_target), 2] i32 out (capacity n_source rows is always enough);  n_pairs [1] i32 out (device).
 */
int d3f_mutual_nn(const float* source, const float* target, int n_source, int n_target, int dim,
                  int32_t* source_arg, int32_t* target_arg, int32_t* pairs, int32_t* n_pairs, d3f_stream stream);

/* ------------------------------------------------------------------------------------------
 * fp32-accurate tensor-core GEMM (3xTF32) with fused epilogue -- the dense contraction behind KPConv
 * (blocks.py:369-380) and the UnaryBlock Linear + bias + LeakyReLU (blocks.py:481-515):
 *   C[M,N] = act( row_scale[m] * sum_k opA(m,k) * k_scale[k] * opB(k,n) + bias[n] )
 *   trans_a: A stored [K,M] else [M,K];  trans_b: B stored [N,K] else [K,N]  (row-major, lda/ldb/ldc in elements)
 *   row_scale / k_scale / bias may be NULL; k_scale needs trans_b == 0; leaky_relu != 0 applies max(v, slope*v).
 */
int d3f_gemm(int trans_a, int trans_b, int M, int N, int K, const float* A, int lda, const float* B, int ldb,
             float* C, int ldc, const float* row_scale, const float* k_scale, const float* bias,
             int leaky_relu, float slope, d3f_stream stream);
int d3f_gemm_prezeroed(int trans_a, int trans_b, int M, int N, int K, const float* A, int lda, const float* B, int ldb,
                       float* C, int ldc, const float* row_scale, const float* k_scale, const float* bias,
                       int leaky_relu, float slope, d3f_stream stream);
/* Deterministic variant (the forward pass): long-K problems are split along K by a rule that depends on K only, the
 * partial tiles go to `workspace` (d3f_gemm_workspace_bytes(M, N, K) bytes, 0 when no split is needed) and are summed
 * in split order before row scale / bias / activation, so a row of C is bit-identical run to run and independent
 * of M.  d3f_gemm instead combines split-K partials with float atomics and never splits a GEMM that has an epilogue.
 * Extra epilogue terms: C = act(... + bias[n] + bias2[n] + residual[m, n]) (bias2 / residual may be NULL): the
 * UnaryBlock's Linear bias + learned bias, and the ResnetBottleneckBlock shortcut add (blocks.py:505-510, :686). */
size_t d3f_gemm_workspace_bytes(int M, int N, int K);
int d3f_gemm_ex(int trans_a, int trans_b, int M, int N, int K, const float* A, int lda, const float* B, int ldb,
                float* C, int ldc, const float* row_scale, const float* k_scale, const float* bias,
                const float* bias2, const float* residual, int ld_residual,
                int leaky_relu, float slope, void* workspace, size_t workspace_bytes, d3f_stream stream);

/* out[0] (device int32) = 1 if a tcgen05 GEMM of this process ever gave up waiting on its mbarrier (the affected output
 * tile is NaN-poisoned); asynchronous device-to-device copy on `stream`, usable inside a CUDA-graph capture. */
int d3f_gemm_status_snapshot(int32_t* out, d3f_stream stream);

/* ------------------------------------------------------------------------------------------
 * Optimiser step on flat buffers.  Replaces torch.optim.SGD.step() as configured by
 * training_3DMatch.py:62-69 (momentum, weight decay, dampening 0) guarded by the reference trainer's
 * "skip the step when a gradient is not finite" loop (trainer.py:104-111), without a host round trip:
 *   g' = g + weight_decay * p;  m = momentum * m + g';  p -= lr[0] * m       (params, momentum_buf updated in place)
 *   params, grads, momentum_buf: n fp32 elements each, 16-byte aligned;  lr: device float (ExponentialLR multiplies it
 *   between epochs, training_3DMatch.py:77-80);  nonfinite_flag: device int32, OR-ed with 1 when check_finite != 0 and
 *   some gradient element is inf / nan; when the flag is non-zero the update is skipped.  The caller clears the flag.
 *   zero_grads != 0: `grads` is cleared in the same pass (the next step's optimizer.zero_grad()).
 * The *_prezeroed variants of d3f_gemm / d3f_colsum / d3f_leaky_backward_colsum are for outputs inside such a cleared
 * gradient buffer: they accumulate (split-K / row-block partial sums, with atomics) without a zero fill of their own.
 */
int d3f_sgd_step(float* params, float* grads, float* momentum_buf, size_t n, const float* lr,
                 float momentum, float weight_decay, int32_t* nonfinite_flag, int check_finite, int zero_grads,
                 d3f_stream stream);

#ifdef __cplusplus
}
#endif
#endif /* D3FEAT_B200_H */
