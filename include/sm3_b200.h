/*
 * sm3_b200.h -- C ABI of the B200-native SM3 contrastive hot path (libsm3_b200.so).
 *
 * Drop-in boundary for the reference's loss path.  The reference (Dylan-H-Wang/skin-sm3) is pure
 * Python/PyTorch; the entry points below are what its modules bind to through ctypes (see
 * INTEGRATION.md for the binding stub).  Each function cites the reference code it replaces
 * (paths relative to the reference checkout).
 *
 * Conventions
 *   - all pointers are DEVICE pointers on the current CUDA device unless the name ends in _host;
 *   - tensors are contiguous row-major; rows 16-byte aligned for the vectorised paths
 *     (unaligned / ragged shapes take a scalar path inside the same kernels, never the CPU);
 *   - the caller owns every buffer including the workspace (size from the *_workspace_bytes query);
 *   - `stream` is a cudaStream_t passed as void*; nothing here synchronises the device;
 *   - return 0 on success, a negative SM3_ERR_* otherwise; text via sm3_last_error() (thread local);
 *   - no C++ exceptions cross this boundary; functions are re-entrant (autograd calls backward from
 *     a worker thread).
 */
#ifndef SM3_B200_H_
#define SM3_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SM3_ABI_VERSION 1

/* dtype codes */
#define SM3_F32 0
#define SM3_F16 1
#define SM3_BF16 2

/* error codes */
#define SM3_OK 0
#define SM3_ERR_SHAPE (-1)     /* bad shape / alignment / null pointer            */
#define SM3_ERR_DTYPE (-2)     /* unsupported dtype or embedding width for `algo` */
#define SM3_ERR_CUDA (-3)      /* CUDA runtime / driver error (see last_error)    */
#define SM3_ERR_WORKSPACE (-4) /* workspace too small                             */

/* algorithm selector for the similarity kernels */
#define SM3_ALGO_AUTO 0
#define SM3_ALGO_SIMT 1 /* fp32 FMA path: exact-fp32 parity, any D <= 256, any dtype      */
#define SM3_ALGO_TC 2   /* TMA + tcgen05/TMEM path: bf16 rows, D in {64,128,192,256}      */

int sm3_version(void);
const char* sm3_last_error(void);
/* 1 if the current device is sm_100 class (tcgen05 path usable), 0 if not, <0 on error */
int sm3_device_supported(void);

/* ------------------------------------------------------------------------------------------------
 * K1  row L2 normalisation      replaces F.normalize(features, dim=1)
 *                               src/models/simclr.py:62 (SimCLR.forward), :138 (V2), :294 (V3/V32)
 * rows [0,rows_a) come from p_a, rows [rows_a, rows_a+rows_b) from p_b (the reference's
 * torch.cat([proj1(f1), proj2(f2)]) simclr.py:293 without the copy); p_b may be NULL.
 *   z[r]        = p[r] / max(||p[r]||_2, eps)         (stored as z_dtype)
 *   inv_norm[r] = 1 / max(||p[r]||_2, eps)            (fp32)
 * ---------------------------------------------------------------------------------------------- */
int sm3_l2norm_fwd(const void* p_a, int64_t rows_a, const void* p_b, int64_t rows_b, int D, int p_dtype,
                   void* z, int z_dtype, float* inv_norm, float eps, void* stream);

/* backward of the above (autograd of simclr.py:294):
 *   dz = scale * sum_{k<n_partials} dz_partials[k]      (fp32 [rows, D] each, partial stride = rows*D)
 *   dp[r] = inv_norm[r] * (dz[r] - z[r] * <z[r], dz[r]>)     (rows on the eps clamp: inv_norm*dz)   */
int sm3_l2norm_bwd(const float* dz_partials, int n_partials, float scale, const void* z, int z_dtype,
                   const float* inv_norm, float eps, void* dp_a, int64_t rows_a, void* dp_b, int64_t rows_b,
                   int D, int dp_dtype, void* stream);

/* ------------------------------------------------------------------------------------------------
 * K2  fused similarity + InfoNCE statistics (never materialises the [M,M] logits)
 *     replaces  matmul(features, features.T) -> mask/gather/cat -> /T      simclr.py:296-320
 *     (= :64-88, :140-164) and the log-softmax half of nn.CrossEntropyLoss  backbone_train.py:531
 *
 * Row-block form (single GPU: n_local == n_global, pair_offset == 0, z_rows == z_cols):
 *   z_cols : [2*n_global, D]  all normalised rows in the reference's order [all first views ; all second]
 *   z_rows : [2*n_local , D]  this rank's rows  [first halves ; second halves]; local row l has global
 *            index  g(l) = pair_offset + l            (l <  n_local)
 *                          n_global + pair_offset + l - n_local   (l >= n_local)
 * Outputs, one per local row i (global index g):
 *   pos[i]     = <z_g, z_pos(g)> * inv_T,       pos(g) = (g + n_global) mod 2*n_global
 *   neg_sum[i] = sum_{j not in {g, pos(g)}} exp((<z_g, z_j> - 1) * inv_T)        (shifted sum)
 *   lse_neg[i] = inv_T + log(neg_sum[i])
 * so that stock CE on logits [pos, lse_neg] with target 0 equals the reference's CE on [M, M-1].
 * ---------------------------------------------------------------------------------------------- */
size_t sm3_infonce_workspace_bytes(int n_local, int n_global, int D, int dtype, int algo, int backward);

int sm3_infonce_fwd(const void* z_rows, const void* z_cols, int n_local, int pair_offset, int n_global, int D,
                    int dtype, float inv_T, float* pos, float* lse_neg, float* neg_sum, void* workspace,
                    size_t workspace_bytes, int algo, void* stream);

/* K3  backward: d/dz_rows of  sum_i (g_pos_i * pos_i + g_lse_i * lse_neg_i)  over ALL global rows,
 *     restricted to this rank's rows (S symmetric => row-local, no column reduce-scatter):
 *       a_j   = g_lse_j / neg_sum_j
 *       H_ij  = exp((s_ij - 1) inv_T) (a_i + a_j)   j not in {i, pos(i)} ;  H_{i,pos(i)} = g_pos_i + g_pos_pos(i)
 *       dz_i  = inv_T * sum_j H_ij z_j
 *     replaces the autograd backward of simclr.py:300-320 + CE (two GEMMs + index_put scatter).
 *     *_rows arrays have 2*n_local entries, *_cols arrays 2*n_global (identical pointers on one GPU).
 *     Output: dz_partials fp32 [n_partials, 2*n_local, D] inside the workspace; the call returns
 *     n_partials (>=1) -- feed it to sm3_l2norm_bwd.  Negative return = error.                       */
int sm3_infonce_bwd(const void* z_rows, const void* z_cols, int n_local, int pair_offset, int n_global, int D,
                    int dtype, float inv_T, const float* g_pos_rows, const float* g_lse_rows,
                    const float* neg_sum_rows, const float* g_pos_cols, const float* g_lse_cols,
                    const float* neg_sum_cols, void* workspace, size_t workspace_bytes, int algo, void* stream);

/* Same as sm3_infonce_bwd with the per-column statistics packed as float4 rows (g_pos, g_lse, neg_sum, unused) in
 * stats_cols[2*n_global, 4] -- the layout the peer-memory exchange below produces.                              */
int sm3_infonce_bwd_packed(const void* z_rows, const void* z_cols, int n_local, int pair_offset, int n_global, int D,
                           int dtype, float inv_T, const float* g_pos_rows, const float* g_lse_rows,
                           const float* neg_sum_rows, const float* stats_cols, void* workspace, size_t workspace_bytes,
                           int algo, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Cross-rank exchange over NVLink peer memory (new capability; the reference has no cross-rank negatives,
 * src/utils/misc.py:629-659 is an unused non-differentiable all_gather).  Every rank stores its 2*n_local rows
 * into EVERY rank's copy of the global buffer at the global row index (order [all first views ; all second]);
 * peers_host = HOST array of `world` device pointers (peer-mapped, e.g. torch symmetric memory buffer_ptrs),
 * including this rank's own.  The caller separates these stores from the readers with a cross-rank barrier.
 *   sm3_peer_scatter_rows : src [2*n_local, row_bytes]  ->  peer[r][2*n_global, row_bytes]   (row_bytes % 16 == 0)
 *   sm3_peer_scatter_stats: (g_pos, g_lse, neg_sum)[2*n_local] -> peer[r][2*n_global] float4 rows
 * ---------------------------------------------------------------------------------------------- */
int sm3_peer_scatter_rows(const void* src, int n_local, int pair_offset, int n_global, int row_bytes,
                          void* const* peers_host, int world, void* stream);
int sm3_peer_scatter_stats(const float* g_pos, const float* g_lse, const float* neg_sum, int n_local, int pair_offset,
                           int n_global, void* const* peers_host, int world, void* stream);
/* NVSwitch multicast forms: multicast_dst is the multicast mapping of the same symmetric buffer (one multimem.st per
 * 16 bytes reaches every rank, so each GPU sends its rows once instead of `world` times).                        */
int sm3_peer_multicast_rows(const void* src, int n_local, int pair_offset, int n_global, int row_bytes,
                            void* multicast_dst, void* stream);
int sm3_peer_multicast_stats(const float* g_pos, const float* g_lse, const float* neg_sum, int n_local, int pair_offset,
                             int n_global, void* multicast_dst, void* stream);

/* Split barrier on a small symmetric flag buffer (>= 64 uint32 per rank, zero-initialised once): `signal` makes this
 * rank's earlier stores visible and writes `epoch` into slot (channel, rank) of every peer's flags; `wait` blocks the
 * stream until all `world` slots of `channel` in MY flags reached `epoch` (epochs only grow; channel in [0,4)).   */
int sm3_peer_signal(void* const* peer_flags_host, int world, int rank, int channel, unsigned epoch, void* stream);
int sm3_peer_wait(const void* my_flags, int world, int channel, unsigned epoch, void* stream);

/* "Remote" halves of K2 / K3 (tcgen05 path, n_local % 128 == 0, world > 1): visit only the column tiles owned by other
 * ranks.  The local block is an ordinary call with n_global = n_local, pair_offset = 0 on the local rows -- it needs no
 * remote data and is enqueued while the exchange is in flight.
 *   fwd_remote : neg_sum = neg_sum_local + remote part ; lse_neg = inv_T + log(neg_sum)   (pos comes from the local call)
 *   bwd_remote : partial dz slabs for the remote columns (returns n_partials); add the local call's slabs.           */
size_t sm3_infonce_remote_workspace_bytes(int n_local, int n_global, int D, int backward);
int sm3_infonce_fwd_remote(const void* z_rows, const void* z_cols, int n_local, int pair_offset, int n_global, int D,
                           int dtype, float inv_T, const float* neg_sum_local, float* pos_unused, float* lse_neg,
                           float* neg_sum, void* workspace, size_t workspace_bytes, void* stream);
int sm3_infonce_bwd_remote_packed(const void* z_rows, const void* z_cols, int n_local, int pair_offset, int n_global,
                                  int D, int dtype, float inv_T, const float* g_pos_rows, const float* g_lse_rows,
                                  const float* neg_sum_rows, const float* stats_cols, void* workspace,
                                  size_t workspace_bytes, void* stream);

/* 1 when sm3_infonce_fwd / sm3_infonce_step run the SYMMETRIC forward for this single-rank shape (rows == columns, bf16,
 * 2 n_pairs % 256 == 0, enough tiles to fill the GPU; SM3_TC_FWD_SYM=0 disables): S = Z Z^T is symmetric, so only the
 * upper-triangular 128 x 128 tiles are computed and a tile's column sums stand in for the transposed tile's row sums --
 * about half the executed MMAs and exponentials for the same result.  *executed_tiles (may be NULL) = S tiles computed. */
int sm3_infonce_fwd_symmetric(int n_pairs, int D, long long* executed_tiles);

/* loss half of nn.CrossEntropyLoss()(logits, 0) on the sufficient statistics, fused with its own
 * gradient (tools/backbone_train.py:101-121):
 *   loss   = scale * sum_i softplus(lse_neg_i - pos_i)             (scale = weight / M_global)
 *   g_lse_i = scale * sigmoid(lse_neg_i - pos_i) ;  g_pos_i = -g_lse_i
 * `loss` is a device scalar; accumulate != 0 adds into it (several terms into one loss).           */
int sm3_infonce_loss(const float* pos, const float* lse_neg, int64_t rows, float scale, float* loss,
                     int accumulate, float* g_pos, float* g_lse, void* stream);

/* ------------------------------------------------------------------------------------------------
 * K4  multi-head softmax cross-entropy, forward + backward in one launch
 *     replaces the 8-head loops  tools/mlc_eval.py:159-162,235-238  tools/backbone_eval.py:102-105
 *     tools/backbone_train.py:178-181  and  tools/mlc_train.py:255-261 (pred / T, ignore_index :381)
 *   logits : [B, sum(class_counts)]   heads concatenated along dim 1 (n = 5,3,2,3,3,3,3,2 -> 24)
 *   labels : [B, H] int64
 *   loss   = (1/H) sum_h w_h * mean_{b: label != ignore} CE(logits_h[b] * inv_T, labels[b,h])
 *   dlogits (may be NULL) = grad_scale * d loss / d logits, same dtype as logits.
 *   class_counts_host / weights_host are HOST arrays of length H (weights may be NULL = all 1).
 *   use_ignore_index == 0 promises no label equals ignore_index (single pass over the data).
 * ---------------------------------------------------------------------------------------------- */
size_t sm3_multihead_ce_workspace_bytes(int64_t B, int H);
int sm3_multihead_ce(const void* logits, int dtype, const int64_t* labels, int64_t B, int H,
                     const int* class_counts_host, const float* weights_host, float inv_T,
                     int use_ignore_index, int64_t ignore_index, float* loss, void* dlogits, float grad_scale,
                     void* workspace, size_t workspace_bytes, void* stream);

/* K5  BCE-with-logits (multi-hot seven-point-checklist targets), forward + backward in one launch.
 *     north_star-named; the reference contains no BCE => parity unpinned, oracle is
 *     torch.nn.functional.binary_cross_entropy_with_logits(reduction="mean").
 *   x, t : [B, C] (t in t_dtype: f32/f16/bf16) ; pos_weight: device fp32 [C] or NULL.            */
size_t sm3_bce_workspace_bytes(int64_t B, int C);
int sm3_bce_logits(const void* x, int x_dtype, const void* t, int t_dtype, const float* pos_weight, int64_t B,
                   int C, float* loss, void* dx, float grad_scale, void* workspace, size_t workspace_bytes,
                   void* stream);

/* ------------------------------------------------------------------------------------------------
 * N1  similarity top-k retrieval   replaces  sim = query @ bank.T ; sim.topk(k)
 *     src/models/evaluator.py:61-63 (KNNOnlineEvaluator.predict)
 *   vals [Bq, k] fp32 descending, idx [Bq, k] int64; ties broken towards the lower bank index.
 *   exclude_self_offset >= 0 masks bank column (exclude_self_offset + q) for query q (retrieval
 *   inside one batch: the reference's "non-self neighbour" test backbone_train.py:103-105); -1 = off.
 * ---------------------------------------------------------------------------------------------- */
int sm3_sim_topk(const void* query, const void* bank, int64_t n_query, int64_t n_bank, int D, int dtype, int k,
                 int64_t exclude_self_offset, float* vals, int64_t* idx, void* stream);
/* Tiled form of the same search (the one the Python front end uses): register-tiled FP32 contraction, threshold-filtered
 * candidate buffers, the bank split across CTAs and a merge kernel; needs a caller-owned 8-byte-aligned workspace of
 * sm3_sim_topk_workspace_bytes().  Identical results (exact top-k, ties -> lower bank index). */
size_t sm3_sim_topk_workspace_bytes(int64_t n_query, int64_t n_bank, int k);
int sm3_sim_topk_ws(const void* query, const void* bank, int64_t n_query, int64_t n_bank, int D, int dtype, int k,
                    int64_t exclude_self_offset, float* vals, int64_t* idx, void* workspace, size_t workspace_bytes,
                    void* stream);
/* Third form: the similarities of up to ~256 MB worth of queries at a time are materialised in the workspace (fp32
 * [chunk, n_bank]) by a register-tiled FP32 contraction over all SMs, then one CTA per query radix-selects the k-th
 * largest key, collects the elements above it plus the lowest-index ties and sorts those k.  Same results.  Workspace:
 * sm3_sim_topk_mat_workspace_bytes(), 16-byte aligned. */
size_t sm3_sim_topk_mat_workspace_bytes(int64_t n_query, int64_t n_bank, int k);
int sm3_sim_topk_mat(const void* query, const void* bank, int64_t n_query, int64_t n_bank, int D, int dtype, int k,
                     int64_t exclude_self_offset, float* vals, int64_t* idx, void* workspace, size_t workspace_bytes,
                     void* stream);

/* ------------------------------------------------------------------------------------------------
 * N3  prototype heads of the multi-label block    replaces tools/mlc_train.py:81-87 (Model.forward)
 *       for i: sa_feats[i] = F.normalize(sa_feats[i], dim=-1)                     (when l2_norm)
 *       preds = [prototypes[i](sa_feats[i % len(sa_feats)]) for i in range(8)]    (bias-free Linears)
 *   feats [Hf, B, D] (Hf = len(sa_feats): 8, or 1 with Identity projectors), W_cat [C, D] fp32 = the prototype weights
 *   concatenated in head order (C = 24), class_slot_host[c] = feature slot class c reads (head(c) % Hf).
 *   fwd: z_out = normalised rows (feats' dtype; untouched when !l2_norm), inv_norm [Hf*B], logits [B, C] fp32 in the
 *        layout sm3_multihead_ce consumes.   bwd: d_feats = d(loss)/d(feats) from dlogits [B, C] (+ d_extra, the gradient
 *        that reached the returned sa_feats, may be NULL); dW_cat = dlogits^T z is a plain GEMM the caller runs.
 *   Shapes: D % 128 == 0 (fp32) / % 256 (16-bit), D <= 512 / 1024, C <= 64 (sm3_proto_heads_supported() == 1).
 * ---------------------------------------------------------------------------------------------- */
int sm3_proto_heads_supported(int D, int C, int dtype);
int sm3_proto_heads_fwd(const void* feats, int dtype, int Hf, int64_t B, int D, const float* W_cat, int C,
                        const int* class_slot_host, int l2_norm, float eps, void* z_out, float* inv_norm, float* logits,
                        void* stream);
int sm3_proto_heads_bwd(const void* z_or_feats, int dtype, int Hf, int64_t B, int D, const float* W_cat, int C,
                        const int* class_slot_host, int l2_norm, const float* inv_norm, const float* dlogits,
                        const void* d_extra, void* d_feats, void* stream);

/* ------------------------------------------------------------------------------------------------
 * N4  DeepCluster spherical k-means of the memory bank    replaces the rank-0 loop of cluster_memory
 *     tools/mlc_train.py:144-176 (torch.mm E step :153, .cpu().numpy() + scipy.sparse + Python loop M step :157-172,
 *     F.normalize :175)
 *   sm3_kmeans_assign : assign[i] = argmax_k <emb_i, centroid_k> (ties -> lower k); when sums/counts are given, also
 *                       sums[k] = sum of the rows assigned to k, counts[k] = how many (fp32), in the SAME pass over the
 *                       bank, deterministically (fixed-order folds, no float atomics).  emb fp32 [n, D], centroids [K, D].
 *   sm3_kmeans_update : centroid_k = sums_k / counts_k where counts_k > 0 (else the old centroid, :173), then every
 *                       centroid is L2-normalised (eps as F.normalize).  A multi-rank caller all-reduces sums / counts
 *                       between the two calls and keeps its bank shard local.
 *   Shapes: fp32, D % 128 == 0, D <= 512, K <= 8 (sm3_kmeans_supported() == 1); the Python front end falls back to
 *   sm3_sim_topk + a one-hot GEMM otherwise.
 * ---------------------------------------------------------------------------------------------- */
int sm3_kmeans_supported(int D, int K, int dtype);
size_t sm3_kmeans_workspace_bytes(int64_t n, int D, int K);
int sm3_kmeans_assign(const float* emb, int64_t n, int D, const float* centroids, int K, int64_t* assign, float* sums,
                      float* counts, void* workspace, size_t workspace_bytes, void* stream);
int sm3_kmeans_update(const float* sums, const float* counts, const float* centroids_old, float* centroids_new, int D,
                      int K, float eps, void* stream);

/* ------------------------------------------------------------------------------------------------
 * N2  fused projector tail    replaces the last two layers of make_projector -- nn.Linear(in_dim, proj_dim, bias=False)
 *     and nn.BatchNorm1d(proj_dim, affine=False), src/models/simclr.py:25-26 -- together with the F.normalize that
 *     follows them (:62, :138, :294), and their autograd backward.
 *   sm3_proj_tail_gemm  : Y [R, D] fp32 = H [R, K] W[D, K]^T on the tensor cores (fp16 / bf16 operands, fp32
 *                         accumulation) and totals[2 D] = per-column (sum Y, sum Y^2) from the same epilogue
 *   (SyncBatchNorm, tools/backbone_train.py:510: the caller all-reduces totals and the row count between the two calls)
 *   sm3_proj_tail_bn_l2 : y^ = (Y - mean) rstd (batch statistics + running-stat update in training mode, running
 *                         statistics in eval mode), z = y^ / max(|y^|, l2_eps) as bf16, inv_norm, mean / rstd saved
 *   sm3_proj_tail_bwd1  : dy^ = inv_norm (dz - z <z, dz>) with dz = sum of K3's partial slabs; totals2[2 D] = column sums
 *                         of dy^ and dy^ * y^        (all-reduced by the caller under SyncBatchNorm)
 *   sm3_proj_tail_bwd2  : dY = rstd (dy^ - totals2[0] / count - y^ totals2[1] / count) in H's dtype; dW = dY^T H and
 *                         dH = dY W are plain GEMMs left to the caller (cuBLAS).
 *   Shapes: K % 64 == 0, D in {64, 128, 192, 256}, 16-byte aligned rows (sm3_proj_tail_supported() == 1).
 * ---------------------------------------------------------------------------------------------- */
int sm3_proj_tail_supported(int K, int D, int dtype);
size_t sm3_proj_tail_workspace_bytes(int64_t R, int D);
int sm3_proj_tail_gemm(const void* h, const void* w, int64_t R, int K, int D, int dtype, float* y, float* totals,
                       void* workspace, size_t workspace_bytes, void* stream);
int sm3_proj_tail_bn_l2(const float* y, int64_t R, int D, const float* totals, float count, float bn_eps, float l2_eps,
                        int training, float momentum, float* running_mean, float* running_var, float* mean_out,
                        float* rstd_out, void* z_bf16, float* inv_norm, void* stream);
int sm3_proj_tail_bwd1(const float* dz_partials, int n_partials, int64_t partial_stride, const void* z_bf16,
                       const float* inv_norm, float l2_eps, const float* y, const float* mean, const float* rstd, int64_t R,
                       int D, float* dyhat, float* totals2, void* workspace, size_t workspace_bytes, void* stream);
int sm3_proj_tail_bwd2(const float* dyhat, const float* y, const float* mean, const float* rstd, const float* totals2,
                       float count, int training, int64_t R, int D, void* dy, int dy_dtype, void* stream);

/* Host-buffer convenience entry (the "plugin call" timed end to end by bench.py): copies p1/p2 from
 * HOST memory, runs normalise -> K2 -> loss -> K3 -> normalise-backward on `stream`, copies the loss
 * and both gradients back to HOST memory and synchronises the stream.  All device scratch comes from
 * `device_scratch` (size from sm3_infonce_host_scratch_bytes).  p/dp dtype = io_dtype.              */
size_t sm3_infonce_host_scratch_bytes(int n_pairs, int D, int io_dtype, int algo);
int sm3_infonce_host(const void* p1_host, const void* p2_host, int n_pairs, int D, int io_dtype,
                     float temperature, float* loss_host, void* dp1_host, void* dp2_host, void* device_scratch,
                     size_t scratch_bytes, int algo, void* stream);

/* Pipelined form of the host-buffer entry (what bench.py's `e2e` times): the handle owns three streams (H2D | kernels |
 * D2H) and `depth` I/O slots inside the caller's `device_scratch`, so the copies of step k+1 and k-1 overlap the
 * kernels of step k.  `submit` enqueues one step and returns a ticket (>= 0) without synchronising, except that it
 * first waits for the ticket `depth` submits ago (its slot is reused); `wait` blocks the host until that ticket's loss
 * and gradients are in the host buffers given to its submit.  Host buffers should be pinned (pageable memory works but
 * serialises the copies) and stay untouched until the ticket has been waited for.  dp1_host/dp2_host may both be NULL
 * (forward only).  One handle = one device (the current device at create) and one submitting thread at a time.      */
#define SM3_PIPE_MAX_DEPTH 4
typedef struct sm3_host_pipe sm3_host_pipe;
size_t sm3_host_pipe_scratch_bytes(int n_pairs, int D, int io_dtype, int algo, int depth);
int sm3_host_pipe_create(sm3_host_pipe** out, int n_pairs, int D, int io_dtype, int algo, int depth,
                         void* device_scratch, size_t scratch_bytes);
int64_t sm3_host_pipe_submit(sm3_host_pipe* pipe, const void* p1_host, const void* p2_host, float temperature,
                             float* loss_host, void* dp1_host, void* dp2_host);
int sm3_host_pipe_wait(sm3_host_pipe* pipe, int64_t ticket);
int sm3_host_pipe_destroy(sm3_host_pipe* pipe);

/* Peer mode of the pipeline (multi-rank jobs): the handle is created for this rank's n_local pairs of a job with n_global
 * pairs, and every submit runs sm3_infonce_step_peer (exchange mode 0, 2, 3 or 4, see below) between the copies, with the
 * symmetric buffers of the slot and the epoch the caller chose for that step (they must follow the same sequence on every
 * rank).  wait / destroy are the ordinary ones.  Gradients are those of sm3_infonce_step_peer (sum over ranks of the
 * per-rank mean losses). */
size_t sm3_host_pipe_peer_scratch_bytes(int n_local, int n_global, int D, int io_dtype, int depth);
int sm3_host_pipe_create_peer(sm3_host_pipe** out, int n_local, int n_global, int D, int io_dtype, int depth,
                              void* device_scratch, size_t scratch_bytes);
int64_t sm3_host_pipe_submit_peer(sm3_host_pipe* pipe, const void* p1_host, const void* p2_host, float temperature,
                                  float* loss_host, void* dp1_host, void* dp2_host, int rank, int world, void* z_cols_mine,
                                  void* const* z_peers_host, void* stats_mine, void* const* stats_peers_host,
                                  void* flags_mine, void* const* flags_peers_host, unsigned epoch, int mode);

/* Device-pointer form of the same fused step: ONE call enqueues normalise -> K2 -> loss -> K3 -> normalise-backward on
 * `stream` (no copies, no synchronisation).  loss = weight * mean-CE (device scalar); dp1/dp2 may both be NULL
 * (forward only).  precision follows `algo` (AUTO = tcgen05 on bf16 rows when D allows, else the fp32 FMA kernels). */
size_t sm3_infonce_step_scratch_bytes(int n_pairs, int D, int io_dtype, int algo);
int sm3_infonce_step(const void* p1, const void* p2, int n_pairs, int D, int io_dtype, float temperature, float weight,
                     float* loss, void* dp1, void* dp2, void* device_scratch, size_t scratch_bytes, int algo,
                     void* stream);

/* Grouped form: `num_terms` independent terms of ONE shape enqueued by one call -- the derm / clinic / cross / cross
 * terms of SimCLRSkinV3 style 0 (tools/backbone_train.py:101-121: loss = derm + clinic + 0.5*cross1 + 0.5*cross2).
 * p1/p2/dp1/dp2 are HOST arrays of `num_terms` device pointers (dp arrays may both be NULL: forward only),
 * weights_host a HOST array (NULL = all 1); loss = sum_t weights[t] * mean-CE_t in one device scalar.  The terms run
 * back to back on `stream` and share one step's scratch (size from sm3_infonce_step_multi_scratch_bytes).           */
size_t sm3_infonce_step_multi_scratch_bytes(int n_pairs, int D, int io_dtype, int algo);
int sm3_infonce_step_multi(int num_terms, const void* const* p1_host_array, const void* const* p2_host_array, int n_pairs,
                           int D, int io_dtype, float temperature, const float* weights_host, float* loss,
                           void* const* dp1_host_array, void* const* dp2_host_array, void* device_scratch,
                           size_t scratch_bytes, int algo, void* stream);

/* Multi-rank form: this rank's 2*n_local rows against all world*n_local pairs, negatives exchanged over NVLink peer
 * memory, the whole step enqueued by ONE call on two streams (exchange on stream_side, kernels on stream_main; the
 * local column block runs while the exchange is in flight).  z_cols_mine / stats_mine / flags_mine are this rank's
 * symmetric buffers ([2*n_global, D] bf16, [2*n_global, 4] fp32, >= 64 uint32 (128 for mode 2)), *_peers_host the HOST arrays of the
 * `world` peer-mapped pointers of the same buffers; `epoch` must grow by one per call and be identical on all ranks.
 * overlap (mode) = 0: exchange, barrier and full-width kernels back to back on stream_main (13 launches);
 *   1: local column block computed while the exchange runs on stream_side (needs n_local % 128 == 0);
 *   2: fused exchange, 5 launches on stream_main: the normalise kernel stores its rows into every rank's z_cols and
 *      signals, the loss kernel does the same for the backward statistics, and K2 / K3 wait for the peers' flags
 *      inside the kernel, visiting this rank's own column tiles first (needs n_local % 128 == 0; flags buffers of
 *      >= 128 uint32, zero-initialised: words [64, 66) are local ticket counters; the stats buffer then holds the
 *      planes a_j = g_lse_j / neg_sum_j at [0, 2*n_global) and g_pos_j at [2*n_global, 4*n_global)).
 *   3: as 2, but the rows are pushed from INSIDE K2: the normalise kernel only fills this rank's own rows, two extra
 *      warps of K2's first-wave CTAs store the rows into the peers' z_cols destination by destination (rank+1 first) and
 *      release one flag per destination, while the tensor-core warps visit the column tiles owner by owner (own columns,
 *      then rank-1's, rank-2's, ...: the order the rows arrive) with one flag wait per source rank -- the whole scatter
 *      hides behind compute.  Uses flag words [66, 85) as tickets; falls back to 2 for problems too small for the
 *      256-row forward kernel.
 *   4: as 2, with the SYMMETRIC forward across ranks (6 launches): S is symmetric, so of the blocks (rows of r, columns
 *      of q) and (rows of q, columns of r) only one is computed.  Rank r takes its own block (upper-triangular tiles), the
 *      full blocks against ranks r+1 ... r+(world-1)/2 and, for even world, half of the block against rank r+world/2 --
 *      world/2 of the world column blocks, the same on every rank --, keeps the row sums and ships the COLUMN sums of the
 *      foreign blocks to their owners (plane 2 of the statistics buffer, [source rank][2*n_local], flag channel 2, ticket
 *      word 88); the loss kernel adds what it received in a fixed order.  Falls back to 2 when n_local % 128 != 0.
 * D in {64,128,192,256}.
 * loss = weight * mean over this rank's rows (DDP convention).                                                     */
size_t sm3_infonce_step_peer_scratch_bytes(int n_local, int n_global, int D);
int sm3_infonce_step_peer(const void* p1, const void* p2, int n_local, int rank, int world, int D, int io_dtype,
                          float temperature, float weight, float* loss, void* dp1, void* dp2, void* z_cols_mine,
                          void* const* z_peers_host, void* stats_mine, void* const* stats_peers_host, void* flags_mine,
                          void* const* flags_peers_host, unsigned epoch, int overlap, void* device_scratch,
                          size_t scratch_bytes, void* stream_main, void* stream_side);

/* out_t[i] = in_t[i] * (*g_device) for `count` (<= 8) tensors of `numel` elements each, one launch: the upstream scaling
 * that autograd applies to the gradients the fused steps computed eagerly (GradScaler factor x loss weight,
 * tools/backbone_train.py:101-125); in/out dtypes may differ (fp32 gradients -> fp16 after the factor). */
int sm3_scale_grads(const void* const* in_host_array, void* const* out_host_array, int count, int64_t numel,
                    int in_dtype, int out_dtype, const float* g_device, void* stream);

/* Per-stage timing of the fused steps on the PRODUCTION path (bench / profiling; not thread safe, one step at a time):
 * while enabled, sm3_infonce_step and sm3_infonce_step_peer (fused mode) record a CUDA event on the launching stream
 * before the first and after every stage; `read` synchronises on the last event of the most recent step and returns the
 * number of stages, ms[k] = duration of stage k; `names` = comma-separated stage names of that step. */
int sm3_stage_timing(int enable);
int sm3_stage_timing_read(float* ms, int capacity);
const char* sm3_stage_timing_names(void);

/* debug / tuning: drop the cached values of the SM3_TC_* environment knobs (SM3_TC_GROUPS, SM3_TC_POLY, SM3_TC_BWD_NS;
 * SM3_TC_FWD_BM, SM3_TC_FWD_SPLITS, SM3_TC_BWD_SPLITS and the HBM-kernel variant selectors are read at every launch)
 * so that a sweep can change them inside one process.  Not a product path.                                       */
void sm3_debug_reload_env(void);

/* debug / bring-up: single-tile tcgen05 probe used by tests/test_umma_probe.py (not a product path).
 *   C[128, n] (fp32) = A[128, k] * B   with the operand sources / layouts selected by `variant`.   */
int sm3_debug_umma_probe(const void* a_bf16, const void* b_bf16, float* c, int n, int k, int variant,
                         void* stream);

/* Host-only enumeration of the tiles the symmetric forward kernels visit (the kernels' own index functions; no GPU
 * needed): per visited tile 5 ints (cta, row pair, GLOBAL column tile, column sums taken, slab row).  world == 1: the
 * single-rank kernel over n_local pairs; world > 1: rank `rank` of exchange mode 4.  tpc_override > 0 forces the tiles per
 * CTA.  Returns the number of records (only the first `cap` are written). */
long long sm3_debug_sym_enumerate(int n_local, int world, int rank, int tpc_override, int* records, long long cap);

/* debug / single-GPU test of mode 4 (symmetric forward across ranks): the caller plays every rank on one device.
 * _forward = rank `rank`'s K2 (flags_mine must hold `epoch` in channel 0 for all sources) + its column-sum push into the
 * per-rank statistics / flag buffers; _fold = that rank's loss kernel without publishing (neg_sum, gradients of the
 * statistics, loss = mean over its rows), to be enqueued after every rank's _forward. */
size_t sm3_debug_mr_workspace(int n_local, int world, int rank);
int sm3_debug_mr_forward(const void* z_local, const void* z_all, int n_local, int world, int rank, int D, float inv_T,
                         void* flags_mine, void* const* stats_peers_host, void* const* flags_peers_host, unsigned epoch,
                         float* pos, void* workspace, size_t workspace_bytes, void* stream);
int sm3_debug_mr_fold(const void* workspace, int n_local, int world, int rank, float inv_T, const float* pos,
                      const void* stats_mine, void* flags_mine, unsigned epoch, float* loss, float* neg_sum, float* g_pos,
                      float* g_lse, float* block_ws, void* stream);

/* debug / single-GPU test of mode 3's forward (owner-ordered column tiles, one flag wait per source rank) with the caller
 * standing in for the peers: z_cols complete, flags[0 .. world) of channel 0 already at `epoch`; nothing is pushed. */
size_t sm3_debug_infonce_fwd_ordered_workspace(int n_local, int world, int D);
int sm3_debug_infonce_fwd_ordered(const void* z_rows, const void* z_cols, int n_local, int rank, int world, int D, float inv_T,
                                  const void* flags, unsigned epoch, float* pos, float* lse_neg, float* neg_sum,
                                  void* workspace, size_t workspace_bytes, void* stream);

/* debug / bring-up microbenchmark (not a product path): dispatch rate of `count` back-to-back tcgen05.mma (M = 128, K = 16,
 * bf16) of width n, A operand from shared memory (0) or TMEM (1), with `ldtm_warps` warps streaming tcgen05.ld meanwhile.
 * out_host[16]: [0] cycles spent issuing, [1] cycles until all completed, [3..] tcgen05.ld round trips of the extra warps. */
int sm3_debug_umma_rate(int n, int a_from_tmem, int count, int ldtm_warps, long long* out_host);

#ifdef __cplusplus
}
#endif
#endif /* SM3_B200_H_ */
