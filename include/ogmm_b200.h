/*
 * ogmm_b200 -- C ABI of the B200-native OGMM registration hot path.
 *
 * One entry point per kernel group of SURVEY.md section 8(b).  Every function
 *   - takes raw DEVICE pointers, element counts, element strides and a CUDA
 *     stream handle (a cudaStream_t passed as void*; NULL = legacy default
 *     stream); no torch types cross this boundary;
 *   - launches on the CALLER's current device and the given stream, never
 *     synchronises the device, holds no global mutable state (the last-error
 *     string is thread-local) -- safe under one-host-thread-per-GPU callers
 *     such as nn.DataParallel (reference train.py:190-192);
 *   - returns OGMM_OK (0) or a negative OGMM_E* code and never throws.
 *
 * All floating point is fp32, all indices are int64, exactly as in the
 * reference.  Tensors named (B,N,C) are LOGICAL shapes; where strides are
 * taken the tensor may be any view (the reference passes transposed views,
 * models/gmmreg.py:26-27, models/dgcnn.py:135).  Strides are in ELEMENTS.
 *
 * The reference is pure Python/PyTorch: it has no FFI of its own.  Each entry
 * point therefore cites the reference FUNCTION (file:line under the reference
 * checkout) whose arithmetic it replaces; INTEGRATION.md shows the ctypes stub
 * and the module-patching a maintainer adds on the reference side.
 */
#ifndef OGMM_B200_H_
#define OGMM_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OGMM_ABI_VERSION 1

#define OGMM_OK            0
#define OGMM_EINVAL       -1   /* bad shape / null pointer / negative size            */
#define OGMM_EUNSUPPORTED -2   /* k, J, N or C outside what the kernels are built for */
#define OGMM_ECUDA        -3   /* launch or runtime failure (see ogmm_last_error)     */
#define OGMM_EWORKSPACE   -4   /* workspace too small                                  */

typedef void* ogmm_stream_t;

/* ABI version of the loaded library (== OGMM_ABI_VERSION). */
int ogmm_version(void);
/* Thread-local description of the last non-zero status returned on this thread. */
const char* ogmm_last_error(void);
/* SM count and compute capability of the current device (any pointer may be NULL). */
int ogmm_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* ---- K1: fused pairwise distance + per-row top-k (+ edge gather) --------------------------
 * Replaces lib/utils.py:12-34 square_distance, :37-44 knn and, when `edge` is non-NULL,
 * :47-66 get_graph_feature (the models/dgcnn.py:135-137 call pair).
 *   src (B,N,C), dst (B,M,C): strided views.   idx_out (B,N,k) int64 contiguous: the k smallest
 *   expanded-form squared distances per row, ascending, ties -> lowest index.
 *   normalize != 0 selects the cosine form 2 - 2 s.d (lib/utils.py:29-30).
 *   edge_out (optional; requires dst == src cloud, i.e. a self graph): (B,N,k,2C) contiguous
 *   memory holding [x_j - x_i ; x_i]; the reference's (B,2C,N,k) result is the permuted view.
 *   dist_out (optional): (B,N,k) the selected distances.
 * C == 3 uses the specialised FP32 kernel; other C go through the generic tiled kernel. */
int ogmm_knn_graph(const float* src, int64_t s_sb, int64_t s_sn, int64_t s_sc,
                   const float* dst, int64_t d_sb, int64_t d_sn, int64_t d_sc,
                   int64_t B, int64_t N, int64_t M, int64_t C, int64_t k, int normalize,
                   int64_t* idx_out, float* dist_out, float* edge_out, ogmm_stream_t stream);

/* Diagnostics of the 3-D selection kernel (knn_select.cu; C == 3, 256 <= M <= 4096, N <= 4096, k <= 24): same idx_out
 * as ogmm_knn_graph plus counters in stats (device int32[16], caller-zeroed): [0] warps that redid their queries
 * exhaustively, [1] warps, [2] sweep steps, [3] steps that merged, [4] insert rounds, [5] sum of the per-warp largest
 * group counts, [7] [8] warps with too many prefix ties / fewer than k groups. */
int ogmm_knn3_select_stats(const float* src, int64_t s_sb, int64_t s_sn, int64_t s_sc,
                           const float* dst, int64_t d_sb, int64_t d_sn, int64_t d_sc,
                           int64_t B, int64_t N, int64_t M, int64_t k,
                           int64_t* idx_out, int32_t* stats, ogmm_stream_t stream);

/* square_distance alone (lib/utils.py:12-34): the dense matrix dist_out (B,N,M) contiguous, same arithmetic as the
 * selection kernels (dist_out of ogmm_knn_graph is a gather of it).  For callers outside the hot path (losses,
 * metrics); the hot path never materialises it. */
int ogmm_square_distance(const float* src, int64_t s_sb, int64_t s_sn, int64_t s_sc,
                         const float* dst, int64_t d_sb, int64_t d_sn, int64_t d_sc,
                         int64_t B, int64_t N, int64_t M, int64_t C, int normalize,
                         float* dist_out, ogmm_stream_t stream);

/* ---- K1w: feature-space kNN on the tensor cores ------------------------------------------------------
 * Same contract as ogmm_knn_graph for 32 <= C <= 256, k <= 32 (ogmm_knn_graph routes such calls here).  The Gram
 * tiles run as tcgen05 TF32 MMAs with TMEM accumulators; a guaranteed superset of the true neighbours is then
 * re-ranked with exact FP32 distances, so idx_out / dist_out are identical to the FP32 kernel's.
 *   fallback_count (optional, device int32, caller-zeroed): diagnostic -- incremented once per selector thread whose
 *   superset buffer filled up and had to be drained early (results are exact either way). */
int ogmm_knn_wide(const float* src, int64_t s_sb, int64_t s_sn, int64_t s_sc,
                  const float* dst, int64_t d_sb, int64_t d_sn, int64_t d_sc,
                  int64_t B, int64_t N, int64_t M, int64_t C, int64_t k, int normalize,
                  int64_t* idx_out, float* dist_out, int32_t* fallback_count, ogmm_stream_t stream);

/* Edge gather alone, for a caller that already holds idx (lib/utils.py:55-66).
 *   x (B,C,N) strided view (strides for b, c, n); idx (B,N,k) int64 contiguous, values in [0,N)
 *   (NOT batch-offset; unlike the reference :57 the index tensor is not modified). */
int ogmm_edge_gather(const float* x, int64_t sb, int64_t sc, int64_t sn, const int64_t* idx,
                     int64_t B, int64_t C, int64_t N, int64_t k, float* edge_out, ogmm_stream_t stream);

/* ---- N3: edge gather fused into the first EdgeConv layer -----------------------------------------
 * Replaces, for inference, models/dgcnn.py:137-141: get_graph_feature (lib/utils.py:47-66) -> conv1 (1x1 Conv2d
 * 6 -> C, no bias) -> bn1 (eval) -> ReLU -> max over k, without materialising the (B,6,N,k) edge tensor.
 *   x (B,3,N) strided view (strides b, c, n); idx (B,N,k) int64 contiguous, values in [0,N);
 *   weight (C,6) contiguous (conv1.weight viewed as (C,6)); scale, shift (C): BatchNorm folded by the caller,
 *   scale = gamma / sqrt(running_var + eps), shift = beta - running_mean * scale.
 *   act_out (optional) (B,C,N,k) contiguous = relu(bn1(conv1(edge))), the tensor conv2 consumes;
 *   max_out (B,C,N) contiguous = its maximum over k (the reference's x1 is max_out viewed as (B,C,N,1)).
 *   k <= 32; the cloud (12 N bytes) must fit shared memory. */
int ogmm_edge_conv_max(const float* x, int64_t x_sb, int64_t x_sc, int64_t x_sn, const int64_t* idx,
                       const float* weight, const float* scale, const float* shift,
                       int64_t B, int64_t N, int64_t k, int64_t C,
                       float* act_out, float* max_out, ogmm_stream_t stream);

/* N3, second half: the k = 5 angle feature of PositionEncoding.forward (models/attn.py:65-73) fused with conv_ang1
 * (Conv2d(1 -> C, bias=False) + BatchNorm2d in eval mode + LeakyReLU(slope)) and the max over the k neighbours:
 *   alpha[b,n,kk] = < normalize(x_j - x_i), normalize(x_i - centroid_b) >,   j = idx[b,n,kk]  (F.normalize, eps 1e-12)
 *   max_out[b,c,n] = max_kk leaky((weight[c] * alpha) * scale[c] + shift[c])
 * x (B,3,N) strided (b,c,n); centroid (B,3) = mean over the points (the caller's torch.mean); idx (B,N,k) int64;
 * weight, scale, shift (C) with BatchNorm folded as in ogmm_edge_conv_max; alpha_out (B,N,k) optional (NULL to skip);
 * max_out (B,C,N).  slope >= 0. */
int ogmm_edge_angle_max(const float* x, int64_t x_sb, int64_t x_sc, int64_t x_sn, const float* centroid,
                        const int64_t* idx, const float* weight, const float* scale, const float* shift, float slope,
                        int64_t B, int64_t N, int64_t k, int64_t C, float* alpha_out, float* max_out,
                        ogmm_stream_t stream);

/* ---- farthest point sampling -----------------------------------------------------------------
 * Replaces lib/utils.py:170-198 farthest_point_sample.  xyz (B,N,3) strided view.
 *   start == NULL: is_center=True (start from the point farthest from the centroid);
 *   otherwise start (B) int64 holds the first index (the reference's torch.randint draw, :190).
 *   ids_out (B,npoint) int64; pts_out (optional) (B,npoint,3) the gathered points (:111-127). */
int ogmm_fps(const float* xyz, int64_t sb, int64_t sn, int64_t sc, int64_t B, int64_t N, int64_t npoint,
             const int64_t* start, int64_t* ids_out, float* pts_out, ogmm_stream_t stream);

/* ---- K2: overlap-guided Sinkhorn k-means (E-step loop + xyz M-step) ---------------------------
 * Replaces lib/utils.py:269-288 wkeans_plus (FPS init :271-272, marginal :276, and per outer
 * iteration cdist :280, sinkhorn :74-108, nan_to_num :282, row normalise :287, gmm_params on xyz
 * :288).  The final feature M-step (:289) is ogmm_gmm_moments_feat.
 *   xyz (B,N,3) strided view; o_scores (B,N) contiguous.
 *   gamma_out (B,N,J), pi_out (B,J), mu_out (B,J,3) contiguous.
 *   The batch-coupled early exit (:99-102) is reproduced exactly: the launch records every cloud's
 *   per-iteration change and its last CTA evaluates the batch means; only when an exit fires does that
 *   CTA tail-launch a redo round from the device (CUDA dynamic parallelism), which re-runs from the first
 *   outer iteration whose inner count changed, and so on until the schedule stands (at most `iters`
 *   rounds).  One host launch, no host synchronisation, no cooperative launch, no grid barrier: no launch
 *   waits on another CTA, so any number of clustering calls may be in flight on one device at once, and
 *   the call can be captured into a CUDA graph.
 *   workspace: device scratch of at least ogmm_sinkhorn_cluster_workspace(...) bytes; contents
 *   need not be initialised.  iters_run_out (optional) (iters) int32: inner iterations per outer. */
int64_t ogmm_sinkhorn_cluster_workspace(int64_t B, int64_t N, int64_t J, int64_t iters, int64_t max_iter);
int ogmm_sinkhorn_cluster(const float* xyz, int64_t sb, int64_t sn, int64_t sc, const float* o_scores,
                          int64_t B, int64_t N, int64_t J, int64_t iters, float tau,
                          float epsilon, float thresh, int64_t max_iter,
                          float* gamma_out, float* pi_out, float* mu_out, int32_t* iters_run_out,
                          void* workspace, int64_t workspace_bytes, ogmm_stream_t stream);

/* Stand-alone log-domain Sinkhorn on a given cost matrix (lib/utils.py:74-108).
 *   cost (B,N,M) contiguous; p (B,N) or NULL (uniform 1/N); q (B,M) or NULL (uniform 1/M).
 *   gamma_out (B,N,M) = exp(K); loss_out (B) = sum gamma*cost per batch element (the reference
 *   returns their mean).  Same early-exit protocol and workspace contract as above. */
int64_t ogmm_sinkhorn_workspace(int64_t B, int64_t N, int64_t M, int64_t max_iter);
int ogmm_sinkhorn(const float* cost, const float* p, const float* q, int64_t B, int64_t N, int64_t M,
                  float epsilon, float thresh, int64_t max_iter, float* gamma_out, float* loss_out,
                  int32_t* iters_run_out, void* workspace, int64_t workspace_bytes, ogmm_stream_t stream);

/* ---- K3: GMM M-step (segmented reductions) -----------------------------------------------------
 * Replaces lib/utils.py:130-149 gmm_params.
 *   gamma (B,N,J) strided view (strides b,n,j); pts (B,N,D) strided view (strides b,n,d).
 *   pi_out (B,J); mu_out (B,J,D); sigma_out (optional) (B,J,D,D) = (sum gamma|x-mu|^2/npi) I.
 * ogmm_gmm_moments is the small-D kernel (D <= 16, e.g. xyz, with optional sigma);
 * ogmm_gmm_moments_feat is the streaming kernel for wide features in their native (B,D,N)
 * layout (pass strides accordingly), reading every feature value from HBM exactly once.
 *   J == 16, N % 4 == 0, unit point stride, 16-byte aligned rows, contiguous gamma: persistent
 *   TMA -> shared memory -> mma.sync 3xTF32 pipeline (FP32-grade accuracy, 1.4e-6 relative).
 *   Any other shape or stride: FP32 FFMA2 kernel.  Environment switches for A/B timing:
 *   OGMM_FEAT_NO_TMA=1 forces the FP32 kernel, OGMM_FEAT_TENSOR=1 selects the tcgen05 variant. */
int ogmm_gmm_moments(const float* gamma, int64_t g_sb, int64_t g_sn, int64_t g_sj,
                     const float* pts, int64_t p_sb, int64_t p_sn, int64_t p_sd,
                     int64_t B, int64_t N, int64_t J, int64_t D,
                     float* pi_out, float* mu_out, float* sigma_out, ogmm_stream_t stream);
int ogmm_gmm_moments_feat(const float* gamma, int64_t g_sb, int64_t g_sn, int64_t g_sj,
                          const float* feats, int64_t f_sb, int64_t f_sn, int64_t f_sd,
                          int64_t B, int64_t N, int64_t J, int64_t D,
                          float* pi_out, float* mu_out, ogmm_stream_t stream);

/* ogmm_gmm_moments_feat with a scratch buffer, for few clouds with very many points (BASELINE.json configs[3]: the
 * cloud x row-block items of 4 clouds leave most SMs idle): the points of a cloud are split over several CTAs whose raw
 * sums a second kernel adds in a fixed order.  ogmm_gmm_moments_feat_workspace returns the bytes needed, 0 when the call
 * gains nothing from the split (the _ws entry point then behaves exactly like ogmm_gmm_moments_feat; workspace may be
 * NULL).  Results agree with the one-pass kernel to FP32 rounding (another summation order). */
int64_t ogmm_gmm_moments_feat_workspace(int64_t B, int64_t N, int64_t J, int64_t D);
int ogmm_gmm_moments_feat_ws(const float* gamma, int64_t g_sb, int64_t g_sn, int64_t g_sj,
                             const float* feats, int64_t f_sb, int64_t f_sn, int64_t f_sd,
                             int64_t B, int64_t N, int64_t J, int64_t D,
                             float* pi_out, float* mu_out, void* workspace, int64_t workspace_bytes,
                             ogmm_stream_t stream);

/* Backward of ogmm_gmm_moments_feat with respect to the features (the reference reaches it
 * through autograd from lib/utils.py:289 in train.py:57-75, with gamma detached at lib/utils.py:286):
 *   grad_feats[b,n,d] = sum_j gamma[b,n,j] * grad_mu[b,j,d] / (pi[b,j] * N + 1e-5)
 *   gamma (B,N,J) strided view; grad_mu (B,J,D) contiguous; pi (B,J) contiguous (the forward's output);
 *   grad_feats: strided (strides b,n,d) -- pass the (B,D,N) buffer's strides to get the gradient in the model's native
 *   feature layout, written with coalesced 16-byte stores. */
int ogmm_gmm_moments_feat_backward(const float* gamma, int64_t g_sb, int64_t g_sn, int64_t g_sj,
                                   const float* grad_mu, const float* pi,
                                   int64_t B, int64_t N, int64_t J, int64_t D,
                                   float* grad_feats, int64_t o_sb, int64_t o_sn, int64_t o_sd, ogmm_stream_t stream);

/* DeepGMR closed-form E-step fused with the M-step (baseline/deepgmr.py:71-74 + lib/utils.py:130-148).
 *   logits (B,J,N) contiguous; pts (B,3,N)-style strided view (strides b,n,d with D == 3).
 *   gamma_out (optional) (B,J,N) softmax over J; pi_out (B,J); mu_out (B,J,3); sigma_out (B,J,3,3). */
int ogmm_softmax_moments(const float* logits, const float* pts, int64_t p_sb, int64_t p_sn, int64_t p_sd,
                         int64_t B, int64_t N, int64_t J,
                         float* gamma_out, float* pi_out, float* mu_out, float* sigma_out,
                         ogmm_stream_t stream);

/* ---- K4: weighted Procrustes and the registration heads ------------------------------------------
 * ogmm_rigid_transform replaces lib/se3.py:256-289 compute_rigid_transformation.
 *   src, corr (B,3,n) strided views (strides b,c,n); weight (B,1,n) as (ptr, stride b, stride n).
 *   rot_out (B,3,3), trans_out (B,3) (the reference's (B,3,1) is a view of it). */
int ogmm_rigid_transform(const float* src, int64_t s_sb, int64_t s_sc, int64_t s_sn,
                         const float* corr, int64_t c_sb, int64_t c_sc, int64_t c_sn,
                         const float* weight, int64_t w_sb, int64_t w_sn,
                         int64_t B, int64_t n, float* rot_out, float* trans_out, ogmm_stream_t stream);

/* ogmm_soft_procrustes replaces models/dgcnn.py:96-115 GMMSVD.forward with is_sk=False (as
 * constructed at models/gmmreg.py:41): cosine similarity (lib/utils.py:222-226), softmax(sim/T),
 * soft correspondences, weights, then the Procrustes above -- one kernel, one CTA per pair.
 *   src_mu (B,Js,3), tgt_mu (B,Jt,3), src_desc (B,Js,D), tgt_desc (B,Jt,D) contiguous.
 *   rot_out (B,3,3), trans_out (B,3), corr_out (B,3,Js); sim_out (optional) (B,Js,Jt). */
int ogmm_soft_procrustes(const float* src_mu, const float* tgt_mu, const float* src_desc, const float* tgt_desc,
                         int64_t B, int64_t Js, int64_t Jt, int64_t D, float temperature,
                         float* rot_out, float* trans_out, float* corr_out, float* sim_out,
                         ogmm_stream_t stream);

/* Backward of the two heads above -- what the reference gets from autograd through torch.svd when train.py:69-75
 * back-propagates the registration loss on (R, t).  Closed-form chain rule in one launch (fp64 3x3 algebra); forward
 * quantities are recomputed from the inputs, so the forward saves nothing.  grad_rot (B,3,3), grad_trans (B,3) and
 * grad_corr (B,3,Js) are the upstream gradients; each may be NULL (= zero).
 *   ogmm_rigid_transform_backward: inputs as ogmm_rigid_transform; grad_src, grad_corr (B,3,n), grad_weight (B,1,n)
 *   contiguous outputs.
 *   ogmm_soft_procrustes_backward: inputs as ogmm_soft_procrustes; grad_src_mu (B,Js,3), grad_tgt_mu (B,Jt,3),
 *   grad_src_desc (B,Js,D), grad_tgt_desc (B,Jt,D) contiguous outputs. */
int ogmm_rigid_transform_backward(const float* src, int64_t s_sb, int64_t s_sc, int64_t s_sn,
                                  const float* corr, int64_t c_sb, int64_t c_sc, int64_t c_sn,
                                  const float* weight, int64_t w_sb, int64_t w_sn, int64_t B, int64_t n,
                                  const float* grad_rot, const float* grad_trans,
                                  float* grad_src, float* grad_corr, float* grad_weight, ogmm_stream_t stream);
int ogmm_soft_procrustes_backward(const float* src_mu, const float* tgt_mu, const float* src_desc,
                                  const float* tgt_desc, int64_t B, int64_t Js, int64_t Jt, int64_t D,
                                  float temperature, const float* grad_rot, const float* grad_trans,
                                  const float* grad_corr, float* grad_src_mu, float* grad_tgt_mu,
                                  float* grad_src_desc, float* grad_tgt_desc, ogmm_stream_t stream);

/* Cosine similarity alone (lib/utils.py:222-226): x (B,N,D), y (B,M,D) contiguous -> (B,N,M).
 * Sized for component descriptors (the GMMSVD head, N = Js, M = Jt): one CTA holds the N x M similarity tile in
 * shared memory, so N * M is limited to about 45,000 entries (e.g. 200 x 200); larger calls return
 * OGMM_EUNSUPPORTED.  (The model's dense N x M point similarity, models/gmmreg.py:75, stays a PyTorch einsum.) */
int ogmm_cos_similarity(const float* x, const float* y, int64_t B, int64_t N, int64_t M, int64_t D,
                        float* sim_out, ogmm_stream_t stream);

/* ogmm_gmm_register replaces baseline/deepgmr.py:17-38 gmm_register.
 *   pi_s (B,J), mu_s (B,J,3), mu_t (B,J,3), sigma_t (B,J,3,3) contiguous -> transform_out (B,4,4). */
int ogmm_gmm_register(const float* pi_s, const float* mu_s, const float* mu_t, const float* sigma_t,
                      int64_t B, int64_t J, float* transform_out, ogmm_stream_t stream);

/* Backward of the DeepGMR path (baseline/deepgmr.py:64-79 under training): the transform's gradient flows through
 * gmm_register into (pi, mu, sigma) and through the xyz moments into gamma = softmax(logits).
 *   ogmm_gmm_register_backward: inputs as ogmm_gmm_register; grad_transform (B,4,4) (the bottom row is ignored);
 *   outputs grad_pi_s (B,J), grad_mu_s, grad_mu_t (B,J,3), grad_sigma_t (B,J,3,3), contiguous.
 *   ogmm_gmm_moments_backward: dL/dgamma of lib/utils.py:130-149 on 3-D points.  pts (B,N,3) strided (b,n,c); pi (B,J),
 *   mu (B,J,3), sigma (B,J,3,3) or NULL as the forward returned them; grad_pi / grad_mu / grad_sigma upstream (each may be
 *   NULL); grad_gamma (B,N,J) written with the given element strides (b,n,j).  J <= 1024. */
int ogmm_gmm_register_backward(const float* pi_s, const float* mu_s, const float* mu_t, const float* sigma_t,
                               int64_t B, int64_t J, const float* grad_transform, float* grad_pi_s, float* grad_mu_s,
                               float* grad_mu_t, float* grad_sigma_t, ogmm_stream_t stream);
int ogmm_gmm_moments_backward(const float* pts, int64_t p_sb, int64_t p_sn, int64_t p_sc, const float* pi,
                              const float* mu, const float* sigma, const float* grad_pi, const float* grad_mu,
                              const float* grad_sigma, int64_t B, int64_t N, int64_t J, float* grad_gamma,
                              int64_t o_sb, int64_t o_sn, int64_t o_sj, ogmm_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* OGMM_B200_H_ */
