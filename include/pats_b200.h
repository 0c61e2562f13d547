/*
 * pats_b200.h -- C ABI of the B200-native (sm_100a) PATS hot path.
 *
 * Drop-in boundary: plain pointers and sizes, no torch types.  Every entry point names the
 * reference interface (zju3dv/pats, file:line) it replaces.  All `*_f32` functions compute in
 * IEEE float32.  Unless the name ends in `_host`, every data pointer is a DEVICE pointer on the
 * current CUDA device and the call only ENQUEUES work on `stream` (a cudaStream_t passed as
 * void*; NULL = legacy default stream) -- no host synchronisation, usable under CUDA graphs.
 * `_host` variants take HOST buffers, copy in, run, copy out and synchronise before returning.
 *
 * Return value: 0 on success, negative PATS_E_* on failure; pats_last_error() returns a
 * thread-local description.  Outputs never alias inputs.
 *
 * There is no CPU implementation behind this header: if the library was built without a kernel
 * for the device found at run time, the calls fail with PATS_E_CUDA.
 */
#ifndef PATS_B200_H
#define PATS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PATS_B200_VERSION 100 /* 0.1.0 */

enum {
    PATS_OK = 0,
    PATS_E_INVALID = -1,     /* bad argument (null pointer, non-positive size, unsupported shape) */
    PATS_E_CUDA = -2,        /* CUDA runtime error (launch failure, no sm_100 device, out of memory) */
    PATS_E_BAD_CROP = -3,    /* tensor_resize: a bound row the reference would reject (library.cpp:55-59 narrow) */
    PATS_E_CAPACITY = -4     /* an output buffer is too small for the result */
};

int pats_version(void);
const char *pats_last_error(void);
/* Number of SMs of the current device, 0 if no CUDA device is usable. */
int pats_sm_count(void);

/* ---------------------------------------------------------------------------------------------
 * Sinkhorn / optimal transport            (replaces models/modules.py:137-182)
 * ------------------------------------------------------------------------------------------- */

/* log_sinkhorn_iterations(Z, log_mu, log_nu, iters)         models/modules.py:137-143
 *   Z [b,M,N], log_mu [b,M], log_nu [b,N]  ->  out [b,M,N]   (row-major, contiguous) */
int pats_log_sinkhorn_iterations_f32(const float *Z, const float *log_mu, const float *log_nu, int b, int M, int N,
                                     int iters, float *out, void *stream);

/* log_optimal_transport(scores, alpha, ns, iters)           models/modules.py:145-162
 *   scores [b,m,n], alpha: DEVICE pointer to one float (the reference passes a 0-dim tensor,
 *   first_layer.py:114), ns [b,1,n]  ->  out [b,m+1,n+1]; dustbin row/column synthesised in-kernel. */
int pats_log_optimal_transport_f32(const float *scores, const float *alpha, const float *ns, int b, int m, int n,
                                   int iters, float *out, void *stream);

/* log_optimal_transport2(scores, one, ns, iters)            models/modules.py:165-182
 *   scores [b,m,n] (dustbin = last row / column), one: DEVICE pointer to one float,
 *   ns [b,1,n-1]  ->  out [b,m,n] (fresh buffer: second_layer.py:108-112 mutates it in place). */
int pats_log_optimal_transport2_f32(const float *scores, const float *one, const float *ns, int b, int m, int n,
                                    int iters, float *out, void *stream);

/* Which kernel family the dispatcher picks for a shape (0 = register-resident warp kernel, 1 = register-resident
 * CTA kernel, 3 = register-resident 8-CTA cluster kernel, 4 = grid-cooperative streaming kernel for plans beyond
 * 512 x 512 (up to 4097 columns), 2 = generic log-domain kernel); for tests and the bench. */
int pats_sinkhorn_kernel_kind(int M, int N);
/* CTAs per problem of the grid-cooperative kernel (0 = automatic: floor(SMs / b), at least one row per warp);
 * tests / A-B timing. */
void pats_sinkhorn_grid_ctas_per_problem(int g);
/* A-B hook of the grid-cooperative kernels, bits: 1 = 16 warps x 1 CTA per SM instead of 8 warps x 2 CTAs per SM (plans up
 * to 1537 columns); 2 = exactly 4096 core columns: one warp per row with the exponentials recomputed instead of four warps
 * per row with the exponentials kept (the default; same results to rounding, 1.7x faster on one 4097 x 4097 plan). */
void pats_sinkhorn_grid_variant(int v);
/* Force the generic log-domain kernel for every shape (tests: exercises the fallback path). */
void pats_sinkhorn_force_generic(int on);
/* Plan hand-over inside the composite calls pats_second_layer_match_f32 / pats_third_layer_match_f32 (default on): the
 * 145 x 145 / 65 x 65 Sinkhorn kernel publishes a per-problem "plan complete" flag and the kernel that consumes the
 * plans is launched with programmatic stream serialization, so it works on finished problems while the solve's last
 * wave is still running.  0 = plain stream order between the two kernels (tests / A-B timing). */
void pats_plan_handover(int on);
/* Launch chaining (default on): the kernels of the path are launched with programmatic stream serialization and begin
 * with griddepcontrol.wait, so the next kernel is resident when its predecessor finishes (ordering unchanged; only the
 * launch gap between small dependent kernels disappears).  0 = plain launches (tests / A-B timing). */
void pats_launch_chaining(int on);
/* Routing of 65 x 65 problems (tests / A-B timing): 0 = two warps per problem (default), 1 = padded 72 x 68 warp
 * kernel, 2 / 3 = one-warp 65 x 65 kernel compiled for 2 / 3 CTAs per SM. */
void pats_sinkhorn_disable_w65(int mode);
/* Routing of 145 x 145 problems (tests / A-B timing): 0 = 8-warp kernel, two CTAs per SM (default), 1 = padded
 * 160 x 160 CTA kernel, 2 = 9-warp kernel. */
void pats_sinkhorn_disable_c145(int mode);
/* Cluster shape for plans up to 320 x 320 (A-B timing): 0 = auto (10 CTAs x 256 threads for batches up to 8 problems, a
 * non-portable cluster size of which a GPC hosts one at a time; 8 CTAs x 256 beyond), 1 = 4 CTAs x 512 threads,
 * 2 = 8 x 256 with a one-hop all-to-all exchange of the column partials, 3 = 8 x 256, 4 = 10 x 256. */
void pats_sinkhorn_cluster_variant(int v);
/* Problems the register-resident kernels handed to the log-domain fallback since the last reset
 * (device counter, read with a synchronising copy; tests / diagnostics only). */
int pats_sinkhorn_fallback_count(int reset);

/* Fixed-point exit of the 65 x 65 kernel (default on).  log_sinkhorn_iterations (models/modules.py:139-142) runs a FIXED
 * number of iterations; once a whole iteration leaves every beta_j bit-identical, all remaining iterations replay it and cannot
 * change the result, so the kernel leaves its loop there.  The output is bit-identical to running all `iters` iterations
 * (tests/test_gpu_ot.py::test_fixed_point_exit_is_bit_identical).  on = 0: always run every iteration.
 * pats_sinkhorn_iterations_skipped: problem-iterations not executed since the last reset (device counter, synchronising read). */
void pats_sinkhorn_fixed_point_exit(int on);

/* Bulk-copy (TMA engine) staging of the 65 x 65 problems: persistent CTAs, one cp.async.bulk of each problem's 16-byte aligned
 * superset into shared memory (completing on an mbarrier), the result formed in place and written back with one bulk store.
 * Bit-identical to the direct kernel and 9 - 12 % faster (profiles/r02_ab_bulk_staging.json), so it is the default; 0 = direct
 * loads (also taken automatically for log_optimal_transport's un-augmented input and for buffers that are not 16-byte aligned). */
void pats_sinkhorn_bulk_staging(int on);
long long pats_sinkhorn_iterations_skipped(int reset);

/* HOST-buffer variants (end-to-end path: H2D, solve, D2H, synchronise). alpha / one by value. */
int pats_log_optimal_transport_f32_host(const float *scores, float alpha, const float *ns, int b, int m, int n,
                                        int iters, float *out);
int pats_log_optimal_transport2_f32_host(const float *scores, float one, const float *ns, int b, int m, int n,
                                         int iters, float *out);

/* ---------------------------------------------------------------------------------------------
 * Patch subdivision                        (replaces setup/library.cpp and utils/utils.py pieces)
 * ------------------------------------------------------------------------------------------- */

/* tensor_resize.tensor_resize(input, bound)                 setup/library.cpp:47-66, :92-93
 *   input [B,C,Hp,Wp] f32, bound [K,5] i64 rows (y0,y1,x0,x1,img*10000+patch)
 *   -> out [K,C,out_h,out_w] f32 (reference: 96x96).  Crop rows [y0,y1), cols [x0,x1], bilinear,
 *   align_corners.  One launch, bounds read on the device.  `bad_rows` (DEVICE int*, may be NULL)
 *   is incremented once per row the reference would reject; such rows are written as zeros. */
int pats_tensor_resize_f32(const float *input, int B, int C, int Hp, int Wp, const int64_t *bound, int K, int out_h,
                           int out_w, float *out, int *bad_rows, void *stream);
/* Same kernel with an explicit rounding recipe for the lerp (0 = bare expression as compiled, 1..9 =
 * (inner, outer) FMA contractions); used to pin bit-exactness against ATen's CUDA kernel.  The shipping
 * recipe is 5, which is bit-identical to torch's upsample_bilinear2d on CUDA. */
int pats_tensor_resize_f32_variant(const float *input, int B, int C, int Hp, int Wp, const int64_t *bound, int K,
                                   int out_h, int out_w, float *out, int *bad_rows, int variant, void *stream);
int pats_tensor_resize_f32_host(const float *input, int B, int C, int Hp, int Wp, const int64_t *bound, int K,
                                int out_h, int out_w, float *out);

/* origin_extract(left, patch_scale, width, height)          utils/utils.py:1300-1318 (non-swap)
 *   left [B,C,ps*(height+2),ps*(width+2)] of `elem` bytes per element
 *   -> out [B,C,height*width,3ps,3ps]; window p=(i,j) starts at (ps*i, ps*j).  Pure copy. */
int pats_origin_extract(const void *left, int elem, int B, int C, int height, int width, int ps, void *out,
                        void *stream);

/* Compute_imgs -- bound / scale arithmetic                   utils/utils.py:1357-1372,1380-1381
 *   x_scale,y_scale [B,n], average_point [B,n,2], n = height*width
 *   -> bound [B,n,4] i64, x_scale_new,y_scale_new [B,n,2], average_new [B,n,2] */
int pats_compute_bounds_f32(const float *x_scale, const float *y_scale, const float *average_point, int B, int height,
                            int width, int ps, int margin, int64_t *bound, float *x_scale_new, float *y_scale_new,
                            float *average_new, void *stream);

/* Compute_imgs -- fused subdivision                          utils/utils.py:1343-1393
 *   left,right [B,H,W,3] (H = ps*height, W = ps*width) of uint8 (elem=1) or f32 (elem=4);
 *   if_nomatching [B,n] u8.  Pads implicitly (zeros), extracts the left 3ps x 3ps windows
 *   (origin_extract) and crop-resizes the right patches (tensor_resize) for the MATCHED patches
 *   only, in row-major mask order:
 *     new_left  [P,3ps,3ps,3] (same element type as left),  new_right [P,3,3ps,3ps] f32
 *     (the reference hands out new_right as the NHWC *view* of this NCHW buffer, utils.py:1385),
 *     bound5 [P,5] i64 (y0,y1,x0,x1,img*10000+patch), x_scale_new,y_scale_new,average_new [B,n,2].
 *   `capacity` = rows available in new_left/new_right/bound5; `count` (DEVICE int*) receives P.
 *   `bad_rows` as in pats_tensor_resize_f32. */
int pats_compute_imgs(const float *x_scale, const float *y_scale, const float *average_point,
                      const uint8_t *if_nomatching, const void *left, const void *right, int elem, int B, int height,
                      int width, int ps, int margin, void *new_left, float *new_right, int64_t *bound5,
                      float *x_scale_new, float *y_scale_new, float *average_new, int capacity, int *count,
                      int *bad_rows, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Area expansion, regrouping, match assembly      (utils/utils.py, models/second_layer.py, third_layer.py)
 * ------------------------------------------------------------------------------------------- */

/* Iterative_expand_matrix(scores_in, scalex, scaley, limitation, ranges, positions, lower_bound, ..., iter_num)
 *                                                           utils/utils.py:1179-1297 (+ Compute_scaling :1321-1340)
 *   scores_in [b,m+1,n+1] = exp(Z), scalex,scaley [b,n], target grid grid_h x grid_w (n = grid_h*grid_w;
 *   `limitation`, `ranges`, `positions` of the reference are functions of the grid and are not passed)
 *   -> whole_cost, core_cost [b,m], average_point [b,m,2], x_scale,y_scale [b,m], bound [b,m,4] i64,
 *      if_nomatching [b,m] u8 (may be NULL; the mask of utils.py:1194). */
int pats_iterative_expand_matrix_f32(const float *scores_in, const float *scalex, const float *scaley, int b, int m,
                                     int grid_h, int grid_w, float lower_bound, int iter_num, float *whole_cost,
                                     float *core_cost, float *average_point, float *x_scale, float *y_scale,
                                     int64_t *bound, uint8_t *if_nomatching, void *stream);

/* est_position's masks                                      first_layer.py:162-167, second_layer.py:244-249
 *   Z [b,M,N] -> nm1 [b,M-1] = (argmax_j Z[i,:] == dust), nm2 [b,N-1] = (argmax_i Z[:,j] == dust) */
int pats_est_nomatching_f32(const float *Z, int b, int M, int N, int dust, uint8_t *nm1, uint8_t *nm2, void *stream);

/* FirstLayer.est_position / SecondLayer.est_position, fused   first_layer.py:159-178, second_layer.py:240-259
 *   Z [b,n+1,n+1] LOG-domain plan (square, n = grid_h*grid_w), scalex,scaley [b,n].  exp() is applied while the
 *   rows are staged (the reference materialises scores.exp()); the two argmax masks and the area expansion run
 *   back to back.  -> trust_score (= whole_cost), x_scale, y_scale [b,n]; average_point [b,n,2];
 *   if_nomatching1, if_nomatching2 [b,n] u8; core_cost [b,n]; bound [b,n,4] i64. */
int pats_est_position_f32(const float *Z, const float *scalex, const float *scaley, int b, int grid_h, int grid_w,
                          float lower_bound, int iter_num, float *trust_score, float *average_point, float *x_scale,
                          float *y_scale, uint8_t *if_nomatching1, uint8_t *if_nomatching2, float *core_cost,
                          int64_t *bound, void *stream);

/* SecondLayer.forward, the matching block                    models/second_layer.py:103-116
 *   scores = log_optimal_transport2(scores, one, ns, iters); scores[:, :, -1] += edge_add; scores[:, -1, :] += edge_add
 *   (edge_add = log 2 outdoor / log 3 indoor, :108-112; the corner gets it twice); est_position(scores, ...).
 *   scores [b,n+1,n+1] (n = grid_h*grid_w; already x0.1), one: DEVICE scalar, ns [b,1,n], scalex,scaley [b,n]
 *   -> Z_out [b,n+1,n+1] (the plan the reference keeps as 'scores') + every output of pats_est_position_f32.
 *   One call so that the two kernels can be handed over problem by problem (see pats_plan_handover). */
int pats_second_layer_match_f32(const float *scores, const float *one, const float *ns, const float *scalex, const float *scaley,
                                int b, int grid_h, int grid_w, int iters, float edge_add, float lower_bound, int iter_num,
                                float *Z_out, float *trust_score, float *average_point, float *x_scale, float *y_scale,
                                uint8_t *if_nomatching1, uint8_t *if_nomatching2, float *core_cost, int64_t *bound, void *stream);

/* SecondLayer.merge_patches_new / merge_patches_old         models/second_layer.py:189-238 / :137-186
 *   trust_score [P,144] f32 and nm_L2 [P,144] u8 are MUTATED in place exactly as the reference mutates its
 *   arguments; nm_L1 [B,hw] u8; scores_back [B,hw,16,9] f64 in/out (carried across chunks for `new`, zeroed on
 *   return for `old`); out [P,144] u8 = if_nomatching per window cell.  workspace: 2*B*hw+1 ints (device).
 *   merge_new: 1 = merge_patches_new, 0 = merge_patches_old; OR-ed with PATS_MERGE_TIE_FIRST to resolve ties of the
 *   reference's unstable `torch.argsort(...)[..., 0]` (:169 / :230) as ATen's CPU kernel does (first minimum) instead of
 *   as its CUDA kernel does (the default: the reference runs on CUDA tensors, evaluate.py:26-28). */
#define PATS_MERGE_TIE_FIRST 2
int pats_merge_patches(int merge_new, float *trust_score, const uint8_t *nm_L1, uint8_t *nm_L2, double *scores_back, int B,
                       int height, int width, int P, uint8_t *out, int *workspace, void *stream);

/* out[i] = torch.argsort(x[i, 0:9])[0]                      the selection rule of second_layer.py:169 / :230 on its own
 *   x [n,9] f64 (device), out [n] i32.  tie_first = 0: ties resolved as ATen's CUDA sort does (the merge default),
 *   1: first minimum (ATen CPU).  Exists so that the tie rule can be checked against the live op. */
int pats_argsort9_first_f64(const double *x, int n, int tie_first, int *out, void *stream);

/* get_result(batch, if_nomatching, average_point, scale, patch_size, left_choice)   utils/utils.py:189-213
 *   two levels, left_choice all true (models/pats.py:75-77).  Level 0: nm0 [B,n0] u8, pt0,sc0 [B,n0,2],
 *   (ps0,h0,w0); level 1: nm1 [P,n1] u8, pt1,sc1 [P,n1,2], (ps1,h1,w1); P = matched level-0 patches.
 *   -> matches_l, matches_r [capacity,2] f32 filled in row-major mask order; *total (DEVICE i64) = Kf.
 *   workspace (device): 8*(P+1) + 4*(2*B*n0 + 1 + P) bytes, 8-byte aligned. */
int pats_get_result_f32(const uint8_t *nm0, const float *pt0, const float *sc0, int B, int ps0, int h0, int w0,
                        const uint8_t *nm1, const float *pt1, const float *sc1, int P, int ps1, int h1, int w1,
                        float *matches_l, float *matches_r, long long capacity, long long *total, void *workspace,
                        void *stream);

/* ThirdLayer.Compute_result + the outdoor label test        models/third_layer.py:184-217, :166-167
 *   scores = exp(Z) [K,65,65], scale_x,scale_y [K,64], p_s,p_t [K,2] i64 (x,y)
 *   -> mkpts0_f, mkpts1_f [K,16,2] f32, if_matching1 [K,16] u8 */
int pats_third_compute_result_f32(const float *scores, const float *scale_x, const float *scale_y, const int64_t *p_s,
                                  const int64_t *p_t, int K, float *mkpts0_f, float *mkpts1_f, uint8_t *if_matching1,
                                  void *stream);

/* Same, taking the LOG-domain plan Z (third_layer.py:158-160: scores = exp(scores_origin) fused into the load). */
int pats_third_result_from_log_f32(const float *Z, const float *scale_x, const float *scale_y, const int64_t *p_s,
                                   const int64_t *p_t, int K, float *mkpts0_f, float *mkpts1_f, uint8_t *if_matching1,
                                   void *stream);

/* ThirdLayer.forward, the matching block                     models/third_layer.py:158-167
 *   scores_origin = log_optimal_transport2(scores, one, ns, iters); Compute_result(exp(scores_origin), ...) + label test.
 *   scores [K,65,65] (already x0.1), ns [K,1,64] -> Z_out [K,65,65], mkpts0_f, mkpts1_f [K,16,2], if_matching1 [K,16] u8.
 *   One call for the same reason as pats_second_layer_match_f32. */
int pats_third_layer_match_f32(const float *scores, const float *one, const float *ns, const float *scale_x, const float *scale_y,
                               const int64_t *p_s, const int64_t *p_t, int K, int iters, float *Z_out, float *mkpts0_f,
                               float *mkpts1_f, uint8_t *if_matching1, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Descriptor correlation feeding the Sinkhorn solves (tcgen05 tensor cores, 3xTF32)
 *   models/first_layer.py:110-114, models/second_layer.py:100-104, models/third_layer.py:156-158
 * ------------------------------------------------------------------------------------------- */

/* out[p,i,j] = scale * sum_k d0[p,k,i] * d1[p,k,j]       torch.einsum('bdn,bdm->bnm', d0, d1) / d**.5 ; 0.1 * scores
 *   d0 [b,d,n], d1 [b,d,m] f32 contiguous (d a multiple of 8), scale = 0.1 / sqrt(d) -> out [b,n,m] f32, the layout
 *   pats_log_optimal_transport*_f32 reads.  FP32-accurate (operands split into two TF32 halves, three MMAs per K step). */
int pats_correlation_f32(const float *d0, const float *d1, int b, int d, int n, int m, float scale, float *out, void *stream);

/* ---------------------------------------------------------------------------------------------
 * The attention network in front of every matching level      models/modules.py:84-134
 *   AttentionalGNN.forward :119-134 (called first_layer.py:106, second_layer.py:93, third_layer.py:148),
 *   AttentionalPropagation :108-117, MultiHeadedAttention :90-106, attention :84-88, MLP :58-69
 * ------------------------------------------------------------------------------------------- */

/* Parameters.  `raw` (DEVICE) holds, per layer and in this order, the reference's tensors as they sit in its state_dict:
 *   attn.proj.0.weight [D,D] .bias [D], attn.proj.1.* , attn.proj.2.* (query, key, value), attn.merge.weight [D,D] .bias [D],
 *   mlp.0.weight [2D,2D] .bias [2D], mlp.1.weight, .bias, .running_mean, .running_var [2D each] (BatchNorm1d), mlp.3.weight [D,2D] .bias [D]
 * = pats_gnn_raw_floats(layers, D) floats.  pats_gnn_pack_f32 writes pats_gnn_packed_floats(layers, D) floats (DEVICE): the
 * query / key / value rows permuted head-major, the merge convolution and the inference-mode BatchNorm folded into the first
 * MLP convolution (FP64 products).  Pack once per set of weights. */
long long pats_gnn_raw_floats(int layers, int D);
long long pats_gnn_packed_floats(int layers, int D);
int pats_gnn_pack_f32(const float *raw, int layers, int D, int heads, float bn_eps, float *packed, void *stream);

/* desc0, desc1 [B,D,N] f32 (DEVICE) -> out0, out1 [B,D,N]: `layers` rounds of desc += mlp(cat(desc, attn(desc, src, src))) with
 *   src = the same set (cross[l] == 0, 'self') or the other one (cross[l] != 0, 'cross'); `cross` is a HOST array of `layers` bytes.
 *   BatchNorm in inference mode (module.eval(); for train() mode see pats_attentional_gnn_train_f32).
 *   `workspace` (DEVICE): at least pats_gnn_workspace_floats(1, D, N) floats; problems are processed in chunks of as many as fit
 *   (28 * N * D floats each).  D a multiple of 8 and of `heads`, the head dimension even; n <= 160 tokens with head dimension <= 96,
 *   or n <= 96 with head dimension <= 32, or any n with head dimension <= 128 (flash-style pass over the keys).
 *   Arithmetic: the 1x1 convolutions on the tcgen05 tensor cores with FP32-class accuracy (3xTF32); pats_gnn_precision(1) selects
 *   single-pass TF32, which is what cuDNN gives the reference's Conv1d on a GPU (torch.backends.cudnn.allow_tf32 defaults to True);
 *   the attention products in FP32 as in the reference. */
long long pats_gnn_workspace_floats(int chunk, int D, int N);
int pats_attentional_gnn_f32(const float *desc0, const float *desc1, int B, int D, int N, const float *packed, const unsigned char *cross,
                             int layers, int heads, float *out0, float *out1, float *workspace, long long workspace_floats, void *stream);
/* The same network with its BatchNorm layers in train() mode (models/pats.py:112-119 keeps the third layer's network in train()
 * when `if_local` is False -- three of the reference's four configurations): each of the two BatchNorm calls of a layer
 * (models/modules.py:131: layer(desc0, src0), layer(desc1, src1)) normalises with the statistics of ITS batch (all tokens of all
 * problems of that side) and updates the running statistics as torch does (momentum, unbiased variance; side 0 first).
 *   packed   from pats_gnn_pack_train_f32 (as pats_gnn_pack_f32 without folding the BatchNorm)
 *   raw      the parameters as for pats_gnn_pack_f32 (gamma / beta are read from here)
 *   running  [layers][2][2D] f32 (DEVICE): running_mean, running_var of every layer -- updated in place
 *   workspace must hold the whole batch: pats_gnn_workspace_floats(B, D, N) floats. */
int pats_gnn_pack_train_f32(const float *raw, int layers, int D, int heads, float *packed, void *stream);
int pats_attentional_gnn_train_f32(const float *desc0, const float *desc1, int B, int D, int N, const float *packed, const float *raw, float *running,
                                   float momentum, float bn_eps, const unsigned char *cross, int layers, int heads, float *out0, float *out1,
                                   float *workspace, long long workspace_floats, void *stream);
void pats_gnn_precision(int passes);
/* A/B switch: 0 = the packed-FP32 (fma.rn.f32x2) generation of the resident-key attention kernels (default), 1 = the first
 *   generation; 2 / 3 = the packed generation at the level-2 shape with 16 query rows x 10 warps / 12 rows x 13 warps instead of
 *   8 x 20 (measured slower).  Same sums in the same order: bit-identical results. */
void pats_gnn_attention_variant(int v);
/* A/B switch: 0 = the TMA-fed, warp-specialised GEMM (operands pre-split into TF32 halves by their producers) in thread-block
 *                 clusters of two CTAs: two token blocks of one output block, each CTA loads half of the weight tile and multicasts
 *                 it into both (default),
 *             1 = the register-staged GEMM (operands split while they are staged),
 *             2 = as 0 without clusters.  Same products, same accumulation order: bit-identical results. */
void pats_gnn_gemm_variant(int v);

/* ---------------------------------------------------------------------------------------------
 * Feature gathers next to the path                  (models/second_layer.py:71-80, models/third_layer.py:119-146)
 * ------------------------------------------------------------------------------------------- */

/* 12x12 grid sampling of the three stem maps (AvgPool2d(2,1,1) of levels 0/1 fused into the gather)
 *   f0 [N,C0,4R,4R], f1 [N,C1,2R,2R], f2 [N,C2,R,R] (R = row_num = 12) -> out [N,C0+C1+C2,R*R]      second_layer.py:71-80 */
int pats_grid_sample12_f32(const float *f0, const float *f1, const float *f2, int N, int C0, int C1, int C2, int row_num,
                           float *out, void *stream);

/* Third-layer 8x8 window unfold + positional add + rubbish token                                 third_layer.py:119-146
 *   feat [P,C,M,M], mkpts_c [K,2] f32 (x,y; snapped to the 4-grid inside; clamp96 = right-image clamp of :127-128),
 *   b_ids [K] f32, kenc [C,64], rubbish [P,C,144], mkpts0_c [K,2] (left points: choose the rubbish token)
 *   -> out [K,C,65].  `bad_index` (DEVICE int*) counts windows whose flat gather index leaves the tensor
 *   (torch.gather would raise). */
int pats_third_unfold_f32(const float *feat, int P, int C, int M, const float *mkpts_c, const float *b_ids, int K, int clamp96,
                          const float *kenc, const float *rubbish, const float *mkpts0_c, float *out, int *bad_index,
                          void *stream);

#ifdef __cplusplus
}
#endif
#endif /* PATS_B200_H */
