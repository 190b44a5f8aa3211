/*
 * unislam_b200.h -- C-ABI of the B200 (sm_100a) replacement for Uni-SLAM's per-frame
 * differentiable-rendering hot path.
 *
 * The reference has no native code of its own: its operator API for this path is the Python
 * surface of the tiny-cuda-nn torch binding plus a handful of PyTorch host functions.  Every
 * entry point below cites the reference interface (file:line under the reference tree) it
 * replaces.  The thin Python layer in uni-slam_b200/ (ctypes + torch.autograd.Function) is the
 * only intended caller; INTEGRATION.md shows the binding a reference maintainer would add.
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer except the usl_grid_t / usl_mlp_t / usl_field_t /
 *     usl_points_t / usl_..._args_t structs is a DEVICE pointer; all float tensors are fp32, contiguous, row-major.
 *   - the caller (PyTorch) owns every buffer; nothing is retained after the call returns.
 *   - all work is enqueued on `stream` (a cudaStream_t passed as void*); no hidden global
 *     state, re-entrant, CUDA-graph capturable (no allocation, no synchronisation inside).
 *   - return 0 on success, non-zero on error; usl_last_error() gives the message
 *     (thread-local).  No C++ exception crosses the boundary.
 *   - "accumulate" outputs (table / decoder / pose gradients) are atomically ADDED into the
 *     caller's buffer, which the caller zero-fills (tcnn does the same memset).
 */
#ifndef UNISLAM_B200_H
#define UNISLAM_B200_H

#include <stdint.h>

#if defined(__GNUC__)
#define USL_API __attribute__((visibility("default")))
#else
#define USL_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define USL_MAX_LEVELS 16
#define USL_FEATS 2       /* n_features_per_level, src/UNISLAM.py:226 */
#define USL_IN 32         /* decoder input width  = 16 levels x 2 (model.c_dim) */
#define USL_HID 16        /* decoder hidden width, src/networks/decoders.py:35 */
#define USL_ACT_NONE 0
#define USL_ACT_TANH 1
#define USL_ACT_SIGMOID 2

typedef void *usl_stream_t; /* cudaStream_t */

/* One level of the multi-resolution hash grid (tcnn GridEncoding offsets table). */
typedef struct usl_level {
    float scale;     /* exp2f(l*log2f(per_level_scale))*base_resolution - 1 */
    uint32_t res;    /* ceilf(scale)+1 */
    uint32_t size;   /* entries in this level */
    uint32_t offset; /* first entry of the level (in entries, x2 floats) */
    uint32_t hashed; /* 1: CoherentPrime hash, 0: dense linear index */
} usl_level_t;

/* tcnn.Encoding(n_input_dims=3, {HashGrid, n_levels, 2, log2_hashmap_size, base_resolution,
 * per_level_scale}, dtype=float) -- src/UNISLAM.py:242-253. Host struct, passed by pointer,
 * copied by value into kernel arguments. */
typedef struct usl_grid {
    int32_t n_levels;
    uint32_t total_entries;
    usl_level_t levels[USL_MAX_LEVELS];
} usl_grid_t;

/* One decoder (src/networks/decoders.py:49-84). Device pointers, row-major (out,in) like
 * nn.Linear.weight.  Variant A (tcnn_network False): n_hidden=2 with biases.  Variant B
 * (tcnn.Network FullyFusedMLP restated in fp32): n_hidden=1, b*=NULL, wo = rows of the padded
 * 16x16 output matrix.  The same struct with writable pointers receives gradients. */
typedef struct usl_mlp {
    float *w1, *b1; /* [16,32], [16] or NULL */
    float *w2, *b2; /* [16,16], [16] or NULL; ignored when n_hidden == 1 */
    float *wo, *bo; /* [n_out,16], [n_out] or NULL */
    int32_t n_hidden; /* 1 or 2 */
    int32_t n_out;    /* 1 (sdf) or 3 (rgb) */
    int32_t out_act;  /* USL_ACT_* */
    int32_t _pad;
} usl_mlp_t;

/* scene_rep + Decoders: (sdf grid, colour grid) and their two decoders, plus the scene bound
 * (src/UNISLAM.py:205-222) used for point normalisation (src/utils/Renderer.py:136-137). */
typedef struct usl_field {
    usl_grid_t grid[2];      /* [0] sdf, [1] colour */
    const float *table[2];   /* fp32 [total_entries*2] */
    usl_mlp_t mlp[2];        /* [0] sdf decoder (n_out 1, tanh), [1] colour decoder (n_out 3, sigmoid) */
    float bound_lo[3], bound_hi[3];
} usl_field_t;

/* scene bound after UNISLAM.load_bound (host struct) */
typedef struct usl_bound {
    float lo[3], hi[3];
} usl_bound_t;

USL_API const char *usl_last_error(void);
USL_API int usl_version(void);

/* tcnn GridEncoding constructor (offset table), run once on the host. */
USL_API int usl_grid_build(int n_levels, int log2_hashmap_size, int base_resolution, double per_level_scale,
                   usl_grid_t *out);

/* ---- B1: tinycudann.Encoding.forward/backward (src/networks/decoders.py:101-103) ---------- */
/* y[n,2L] = HashGrid(x[n,3]) */
USL_API int usl_grid_encode_fwd(const usl_grid_t *g, const float *params, const float *x, int64_t n, float *y,
                        usl_stream_t stream);
/* grad_params[total*2] += scatter(w * dy)   (tcnn kernel_grid_backward) */
USL_API int usl_grid_encode_bwd_params(const usl_grid_t *g, const float *x, const float *dy, int64_t n,
                               float *grad_params, usl_stream_t stream);
/* dx[n,3] = dy . d y/d x                    (tcnn kernel_grid_backward_input) */
USL_API int usl_grid_encode_bwd_input(const usl_grid_t *g, const float *params, const float *x, const float *dy,
                              int64_t n, float *dx, usl_stream_t stream);
/* parity inspection: idx[n,L,8] = within-level entry index of every corner (bit-exact gate) */
USL_API int usl_grid_corner_indices(const usl_grid_t *g, const float *x, int64_t n, uint32_t *idx,
                            usl_stream_t stream);

/* ---- B2: tinycudann.Network / nn.Linear decoder stacks (decoders.py:107-155) -------------- */
USL_API int usl_mlp_fwd(const usl_mlp_t *m, const float *h, int64_t n, float *out, usl_stream_t stream);
/* grads ADDED into gm->*; dh[n,32] may be NULL */
USL_API int usl_mlp_bwd(const usl_mlp_t *m, const usl_mlp_t *gm, const float *h, const float *out,
                const float *dout, int64_t n, float *dh, usl_stream_t stream);

/* ---- B5: ray generation (src/common.py:95-180, 210-228) ----------------------------------- */
/* get_samples_all: gather-then-rotate from K stored keyframe pixel subsets.
 * c2ws[K,4,4], depths[K,P], colors[K,P,3], dirs_cam[K,P,3], indices[K*n] (int64 torch.randint draw)
 * -> rays_o/rays_d[K*n,3], gt_depth[K*n], gt_color[K*n,3], dirs_out[K*n,3] (camera-frame dir kept for
 * the pose gradient), frame_id[K*n] (int32, + frame_base). */
USL_API int usl_sample_keyframe_rays(const float *c2ws, const float *depths, const float *colors,
                             const float *dirs_cam, const int64_t *indices, int K, int64_t P, int n,
                             int frame_base, float *rays_o, float *rays_d, float *gt_depth,
                             float *gt_color, float *dirs_out, int32_t *frame_id, usl_stream_t stream);
/* get_samples (tracker): indices[n] into the (H1-H0)x(W1-W0) window of one frame. */
USL_API int usl_sample_window_rays(const float *c2w, const float *depth, const float *color, int H, int W,
                           int H0, int H1, int W0, int W1, float fx, float fy, float cx, float cy,
                           const int64_t *indices, int64_t n, float *rays_o, float *rays_d,
                           float *gt_depth, float *gt_color, float *dirs_out, usl_stream_t stream);
/* get_rays: all H*W rays of one camera. */
USL_API int usl_image_rays(const float *c2w, int H, int W, float fx, float fy, float cx, float cy,
                   float *rays_o, float *rays_d, usl_stream_t stream);

/* ---- f2: keyframe store + co-visibility (src/Mapper.py:177-236, 528-541) ---------------------------------------------------
 * A keyframe is stored as a pixel subset: gather colour[H*W,3] / depth[H*W] / camera dirs[H*W,3] at indices[P]
 * (torch.randperm(H*W)[:P]) into the store's rows in one launch (idx_row nullable). */
USL_API int usl_keyframe_insert(const float *color_img, const float *depth_img, const float *dirs_cam, const int64_t *indices,
                                int64_t P, float *color_row, float *depth_row, float *dirs_row, int64_t *idx_row,
                                usl_stream_t stream);
/* keyframe_selection_LC's overlap measure: percent_inside[k] = fraction of the points sampled in [0.8 d, d + 0.5] along the n rays
 * with sensor depth (num_samples each, torch.linspace) that project inside keyframe k's image minus `edge` pixels, in front of
 * the camera (c2ws[K,4,4] estimated poses; OpenGL convention, x flipped as in Mapper.py:222). */
USL_API int usl_keyframe_covisibility(const float *rays_o, const float *rays_d, const float *gt_depth, int64_t n, int num_samples,
                                      const float *c2ws, int K, int H, int W, float fx, float fy, float cx, float cy,
                                      float edge, float *percent_inside, usl_stream_t stream);

/* ---- a-4: ray/bbox exit distance (Mapper.py:396-401, Tracker.py:177-183) ------------------- */
/* t_exit[n]; valid[n] = t_exit >= gt_depth (&& gt_depth > 0 when require_depth) */
USL_API int usl_bbox_prefilter(const float *rays_o, const float *rays_d, const float *gt_depth, int64_t n,
                       const usl_bound_t *bound, int require_depth, float *t_exit, uint8_t *valid,
                       usl_stream_t stream);

/* ---- a-5/a-6: z sampling (src/utils/Renderer.py:42-57,77-130; common.py:49-85) ------------- */
typedef struct usl_zsample_args {
    int32_t n_stratified, n_importance;
    float c_surf_lo;   /* (float)(1.5*truncation) */
    float c_surf_span; /* (float)(3*truncation)   */
    const float *t_uni;  /* torch.linspace(0,1,n_stratified) on device */
    const float *t_surf; /* torch.linspace(0,1,n_importance) on device */
} usl_zsample_args_t;
/* rays with gt_depth>0: z[r,:] = perturb(sort(cat(free, surface))). t_rand row = row_map ? row_map[r] : r;
 * t_rand NULL => perturb off. Rays with valid==0 or gt<=0 are left untouched. */
USL_API int usl_zsample_depth(const usl_zsample_args_t *a, const float *gt_depth, const uint8_t *valid,
                      const float *t_rand, const int32_t *row_map, int64_t n_rays, float *z,
                      usl_stream_t stream);
/* rays with gt_depth==0: uniform to bbox exit + importance resampling from an SDF-only query
 * (the [-1,1]-then-clamp normalisation quirk kept). pdf_inds (nullable): searchsorted indices. */
USL_API int usl_zsample_nodepth(const usl_zsample_args_t *a, const usl_field_t *f, const float *beta,
                        const float *rays_o, const float *rays_d, const float *gt_depth,
                        const uint8_t *valid, const float *t_rand_uni, const float *u_pdf,
                        const int32_t *row_map, int64_t n_rays, float *z, int64_t *pdf_inds,
                        usl_stream_t stream);

/* ---- fused ray set-up: a-3 + a-4 + a-5 in ONE launch (and a-10's pose -> matrix when poses are given) ---------- */
typedef struct usl_ray_batch {          /* one get_samples_all call (src/Mapper.py:379-393) */
    const float *c2ws;                  /* [K,4,4] camera matrices, ignored when cam_poses != NULL */
    const float *depths, *colors, *dirs_cam; /* [K,P], [K,P,3], [K,P,3] stored keyframe pixel subsets */
    const int64_t *indices;             /* [K*n] torch.randint(P, (n*K,)) */
    int64_t P;
    int32_t K, n, frame_base, _pad;
} usl_ray_batch_t;
typedef struct usl_ray_setup {
    int32_t mode;                       /* 0: keyframe batches (mapper), 1: image window (tracker), 2: n_rays consecutive
                                         * pixels of the full frame from pixel_begin (render_img, src/utils/Renderer.py:160-223:
                                         * get_rays over the image, no bbox prefilter, valid = 1) */
    int32_t n_batches;
    usl_ray_batch_t batch[2];
    const float *depth_img, *color_img; /* mode 1: [H,W], [H,W,3] */
    const int64_t *win_indices;         /* mode 1: [n_rays] torch.randint((H1-H0)*(W1-W0), (n,)) */
    int32_t H, W, H0, H1, W0, W1;
    float fx, fy, cx, cy;
    const float *c2w;                   /* mode 1: [4,4] when cam_poses == NULL */
    const float *cam_poses;             /* [K-1,7] (mode 0: frame f>0 uses pose f-1, frame 0 uses c2w_fixed) or [1,7] (mode 1) */
    const float *c2w_fixed;             /* [4,4] */
    usl_bound_t bound;
    int32_t require_depth;              /* tracker: also drop rays with gt_depth <= 0 */
    int32_t _pad2;
    usl_zsample_args_t zs;
    const float *t_rand;                /* [n_rays,S] or NULL (perturb off) */
    int64_t n_rays;
    float *rays_o, *rays_d, *gt_depth, *gt_color, *dirs_out; /* outputs */
    int32_t *frame_id;                  /* nullable */
    uint8_t *valid;
    float *z;                           /* [n_rays,S]; rows of depth-less rays are left for usl_zsample_nodepth */
    int64_t pixel_begin;                /* mode 2: row-major index (j*W + i) of the first pixel; gt_color / dirs_out nullable */
    int64_t ray_offset;                 /* mode 0: first GLOBAL ray slot of this call (multi-GPU: the batch is split contiguously
                                         * across ranks after sampling); indices / t_rand are indexed by global slot,
                                         * outputs by local slot 0..n_rays-1 */
} usl_ray_setup_t;
USL_API int usl_ray_setup(const usl_ray_setup_t *a, usl_stream_t stream);

/* ---- B3/B4: field query (Decoders.forward, decoders.py:182-205, with Renderer.py:132-139) -- */
typedef struct usl_points {
    const float *x;       /* [n,3] normalised coordinates (un-clamped), or NULL to build from rays */
    const float *rays_o;  /* [R,3] */
    const float *rays_d;  /* [R,3] */
    const float *z;       /* [R,S] */
    const uint8_t *valid; /* [R] or NULL */
    int32_t S;
    int32_t sample_major; /* built from rays only. 0: thread i = (ray i / S, sample i % S). 1: thread i = (ray i % R, sample i / R),
                           * R = n / S -- a warp then holds the SAME sample index of 32 consecutive rays; when consecutive rays
                           * are neighbouring pixels (whole-frame rendering) its points are centimetres apart and share cells.
                           * Outputs are written in ray-major order either way. */
    int64_t n;            /* number of points (= R*S when built from rays) */
} usl_points_t;
/* raw[n,4] = (r,g,b,sdf). feat (nullable): activation stash for the backward pass, usl_field_stash_floats(n) = n*(2*48+3)
 * floats: interpolated features [2][L][n][2], the hidden pre-activations [2][16][n], then the clamped coordinates [3][n]
 * (x0 = -1 marks a filtered point: usl_field_bwd needs neither the rays nor z again).  Every row is contiguous over the
 * points, so the backward fetches a tile of it with a handful of bulk copies.  jac (nullable): [12,n] component-major
 * d raw / d x (component o*3+d, rows o = r,g,b,sdf; clamp-gated). */
USL_API int usl_field_stash_floats(int64_t n_points, int64_t *n_floats);
USL_API int usl_field_fwd(const usl_field_t *f, const usl_points_t *p, float *raw, float *feat, float *jac,
                  usl_stream_t stream);
/* usl_field_fwd with the Jacobian's tangent contraction on the 5th-generation tensor cores (tcgen05.mma kind::tf32, TMEM
 * accumulators; csrc/field_tc.cu).  Same arguments and outputs (raw bit-identical, jac within the gradient tolerance).  Measured
 * slower than the CUDA-core kernel on the mapping workload (the kernel is bound by the gather side), hence a separate entry point. */
USL_API int usl_field_fwd_tc(const usl_field_t *f, const usl_points_t *p, float *raw, float *feat, float *jac,
                     usl_stream_t stream);
/* d_raw[n,4] -> table gradients (scatter), decoder gradients (gm[2], may be NULL to skip).  Of `p` only n is read (the
 * points themselves come from the stash).  Persistent kernel: stash / d_raw / raw tiles arrive by cp.async.bulk when
 * n % 4 == 0 and raw, feat, d_raw are 16-byte aligned, by ordinary loads otherwise (same results).
 * scratch (nullable): zero-filled workspace of usl_field_bwd_scratch_floats() floats holding private copies of the
 * small coarse levels (L2 atomics serialise on small tables) and the work-queue counters of the persistent kernel; it
 * is folded into the gradient tables before return and must be zero-filled again by the caller before the next call
 * (NULL: no replicas, static tile assignment). grid_mask: 1 = sdf grid only, 2 = colour grid
 * only, 3 = both (one launch); the halves are independent, so a caller can all-reduce one while the other runs.
 * | 4 (USL_BWD_LEAVE_ROOM): launch one resident CTA per SM fewer than fit, so that a small concurrent kernel (the gradient
 * exchange of the other half) finds registers on every SM whatever the launch order. */
#define USL_BWD_LEAVE_ROOM 4
USL_API int usl_field_bwd_scratch_floats(const usl_field_t *f, int64_t *n_floats);
USL_API int usl_field_bwd(const usl_field_t *f, const usl_points_t *p, const float *raw, const float *feat,
                  const float *d_raw, float *grad_table_sdf, float *grad_table_rgb,
                  const usl_mlp_t *gm, float *scratch, int grid_mask, usl_stream_t stream);
/* SDF channel only, forward only (Renderer.py:121, Mesher.py:134-166). out_of_bound_value used
 * when mask_bound!=0 and the un-normalised point lies outside the open bound. */
USL_API int usl_field_sdf(const usl_field_t *f, const usl_points_t *p, float *sdf, usl_stream_t stream);

/* ---- a-8: SDF->weight compositing (Renderer.py:140-158), one warp per ray ------------------ */
/* outputs: term[R], pixel_unc[R], depth[R], rgb[R,3], depth_unc[R]; weights (nullable) [R,S] */
USL_API int usl_composite_fwd(const float *raw, const float *z, const float *beta, const uint8_t *valid,
                      int64_t R, int S, float *term, float *pixel_unc, float *depth, float *rgb,
                      float *depth_unc, float *weights, usl_stream_t stream);
/* upstream grads (each nullable = zero): g_term[R], g_punc[R], g_depth[R], g_rgb[R,3], g_dunc[R],
 * g_sdf[R,S].  Outputs: d_raw[R,S,4]; d_beta[1] (ADDED); and, when jac!=NULL ([12,R*S]), d_rays_o/d_rays_d[R,3]
 * = sum_s (d_raw . jac) chained through the normalisation (extent = hi-lo). */
USL_API int usl_composite_bwd(const float *raw, const float *z, const float *beta, const uint8_t *valid,
                      int64_t R, int S, const float *g_term, const float *g_punc, const float *g_depth,
                      const float *g_rgb, const float *g_dunc, const float *g_sdf, const float *jac,
                      const usl_bound_t *bound, float *d_raw, float *d_beta, float *d_rays_o,
                      float *d_rays_d, usl_stream_t stream);

/* usl_composite_bwd with the loss gradient computed in place (== usl_loss_bwd + usl_composite_bwd, no g_sdf round trip)
 * and loss[0] written by the first CTA (== usl_loss_finalize). Declared after usl_loss_args_t below. */

/* ---- a-9: masks + losses (Mapper.py:141-175,412-430; Tracker.py:113-147,208-228) ----------- */
typedef struct usl_loss_args {
    float truncation;
    float truncation_center; /* (float)(0.4*truncation), product taken in double by the caller (Mapper.py:160-161) */
    float w_sdf_fs, w_sdf_center, w_sdf_tail, w_depth, w_color;
    int32_t mode;      /* 0 mapping ('original'), 1 tracking ('original': 10x-median mask, masked colour),
                        * 2 'no_mask' of either (every term over every valid ray; Mapper.py:432-440, Tracker.py:230-238) */
} usl_loss_args_t;
#define USL_LOSS_SLOTS 16
/* acc[USL_LOSS_SLOTS] (zeroed by caller; ADDED): 0 fs_sum 1 center_sum 2 tail_sum 3 depth_sum 4 color_sum
 * 5 n_front 6 n_center 7 n_tail 8 n_mask 9 n_rays(valid) 10 n_color_terms 11 sum_pixel_unc.
 * tracking mode reads median[0] (device) for the depth-error mask. mask_out[R] (nullable). */
USL_API int usl_loss_fwd(const usl_loss_args_t *a, const float *raw, const float *z, const float *gt_depth,
                 const float *gt_color, const uint8_t *valid, const float *pixel_unc, const float *depth,
                 const float *rgb, const float *median, int64_t R, int S, float *acc, uint8_t *mask_out,
                 usl_stream_t stream);
/* loss[0] = weighted sum of means from acc (NaN when a mask is empty, as torch.mean of empty) */
USL_API int usl_loss_finalize(const usl_loss_args_t *a, const float *acc, float *loss, usl_stream_t stream);
/* upstream gradients of the loss wrt the render outputs, scaled by g_loss[0] (device, NULL = 1):
 * g_depth[R], g_rgb[R,3], g_sdf[R,S]. */
USL_API int usl_loss_bwd(const usl_loss_args_t *a, const float *raw, const float *z, const float *gt_depth,
                 const float *gt_color, const uint8_t *valid, const uint8_t *mask, const float *depth,
                 const float *rgb, const float *acc, const float *g_loss, int64_t R, int S,
                 float *g_depth, float *g_rgb, float *g_sdf, usl_stream_t stream);
/* usl_composite_fwd + usl_loss_fwd in one launch for the modes whose ray mask needs no global quantity (0 and 2; mode 1 is
 * refused: the tracker's mask needs the median over all rays first).  Same outputs as the two calls (Renderer.py:140-158 +
 * Mapper.py:141-175,412-430); acc is ADDED to. */
USL_API int usl_composite_loss_fwd(const usl_loss_args_t *a, const float *raw, const float *z, const float *beta,
                                   const uint8_t *valid, int64_t R, int S, const float *gt_depth, const float *gt_color,
                                   float *term, float *pixel_unc, float *depth, float *rgb, float *depth_unc, float *acc,
                                   uint8_t *mask_out, usl_stream_t stream);
USL_API int usl_composite_loss_bwd(const usl_loss_args_t *a, const float *raw, const float *z, const float *beta,
                                   const uint8_t *valid, const uint8_t *mask, int64_t R, int S, const float *gt_depth,
                                   const float *gt_color, const float *depth, const float *rgb, const float *acc,
                                   const float *g_loss, const float *jac, const usl_bound_t *bound, float *d_raw,
                                   float *d_beta, float *d_rays_o, float *d_rays_d, float *loss, usl_stream_t stream);
/* torch.median (lower middle) of |gt - depth| over valid rays: median[0]. workspace: >= R floats. */
USL_API int usl_depth_error_median(const float *gt_depth, const float *depth, const uint8_t *valid, int64_t R,
                           float *workspace, float *median, usl_stream_t stream);

/* ---- a-10: pose gradient (common.py:196-208 cam_pose_to_matrix; pytorch3d quaternion_to_matrix) */
/* d_c2w[K,12] (3x4 row-major, ADDED) from per-ray d_rays_o / d_rays_d and the camera-frame dirs. */
USL_API int usl_pose_reduce(const float *d_rays_o, const float *d_rays_d, const float *dirs_cam,
                    const int32_t *frame_id, const uint8_t *valid, int64_t n, int K, float *d_c2w,
                    usl_stream_t stream);
/* pose[K,7]=[qw,qx,qy,qz,tx,ty,tz] -> c2w[K,16] */
USL_API int usl_pose_to_matrix(const float *pose, int K, float *c2w, usl_stream_t stream);
/* d_pose[K,7] = chain of d_c2w[K,12] through quaternion_to_matrix (un-normalised quaternion). */
USL_API int usl_pose_matrix_bwd(const float *pose, const float *d_c2w, int K, float *d_pose, usl_stream_t stream);

/* Tracker.py:346-348 on the device (no .item() sync): if (loss[0] < best_loss[0]) { best_loss[0] = loss[0];
 * best_pose[0..7) = cam_pose[0..7) }  -- "best pose = the pose at which the minimal loss was evaluated". */
USL_API int usl_track_keep_best(const float *loss, const float *cam_pose, float *best_loss, float *best_pose,
                                usl_stream_t stream);

/* ---- a-11: dense SDF query for meshing (src/utils/Mesher.py:134-195) ----------------------- */
/* points generated in-kernel from the per-axis coordinate arrays ax/ay/az (np.linspace -> fp32),
 * ordering of torch.meshgrid(indexing='xy') flattened: idx = (iy*nx + ix)*nz + iz for iy in
 * [y_begin,y_end). SDF = -1 outside the open bound. out[(y_end-y_begin)*nx*nz]. */
USL_API int usl_sdf_query_grid(const usl_field_t *f, const float *ax, const float *ay, const float *az, int nx,
                       int ny, int nz, int y_begin, int y_end, float *out, usl_stream_t stream);

/* ---- a-12 / f1: optimiser step (torch.optim.Adam of src/Mapper.py:358-364,445, src/Tracker.py:324-329,242) ----- */
#define USL_ADAM_MAX_GROUPS 24
typedef struct usl_adam_group {
    float *param, *grad, *exp_avg, *exp_avg_sq; /* device, n floats each */
    int64_t n;
    float lr, beta1, beta2, eps;
    int64_t step;                               /* this tensor's own 1-based step count (torch keeps state['step'] per parameter);
                                                 * 0 = use the call's `step` / `step_dev` */
} usl_adam_group_t;
/* One fused launch over all groups; same update rule as torch.optim.Adam (no weight decay / amsgrad).
 * step: 1-based step count of this update (bias correction); step_dev (nullable): device int64 holding it instead
 * (CUDA-graph replay). zero_grad != 0 also clears the gradients (replaces optimizer.zero_grad + tcnn's memset). */
USL_API int usl_adam_step(const usl_adam_group_t *groups, int n_groups, int64_t step, const int64_t *step_dev,
                          int zero_grad, usl_stream_t stream);

/* ---- f4: marching cubes on the device-resident SDF volume (src/utils/Mesher.py:230-258) ----------------------------------
 * Replaces the D2H copy of the volume + skimage.measure.marching_cubes(volume[x,y,z], level, spacing) + the origin shift.
 * A vertex on every grid edge whose end values straddle the level (inside = value < level) at the linear interpolation
 * t = (level - v0)/(v1 - v0); indexed triangles from a 256-case table (csrc/mc_tables.h, generated; crack-free, differs from
 * skimage's Lewiner tables only in how ambiguous configurations are triangulated; normals point towards larger values).
 * Works on one y-slab of the volume as usl_sdf_query_grid wrote it: vol[(iy*nx + ix)*nz + iz], iy in [0, rows); a slab that is
 * not the last carries one halo row (rows = own_rows + 1).  Call order: usl_mc_classify -> usl_scan_u8(pflags, popcount=1) ->
 * usl_scan_u8(ctri, popcount=0) -> allocate verts[V,3] / faces[T,3] -> usl_mc_emit. */
typedef struct usl_mc_args {
    const float *vol;
    int32_t nx, nz, rows, own_rows;
    int32_t y_begin;                  /* global row index of the slab's first row */
    float level;
    float origin[3], spacing[3];      /* world position of grid index (0,0,0) of the whole volume, grid spacing */
    uint8_t *pflags;                  /* [rows*nx*nz] out (classify), in (emit): bit a = a vertex on the edge towards +axis a */
    uint8_t *ctri;                    /* [rows*nx*nz] out (classify), in (emit): triangles of the cell rooted at this point */
    const uint32_t *voff, *toff;      /* emit: exclusive prefix sums of popcount(pflags) / ctri */
    float *verts;                     /* emit: [V,3] world coordinates */
    int64_t *vkeys;                   /* emit, nullable: [V] 3*(global point index) + axis -- what a multi-GPU merge welds by */
    int32_t *faces;                   /* emit: [T,3] indices into this slab's vertex array */
} usl_mc_args_t;
USL_API int usl_mc_classify(const usl_mc_args_t *a, usl_stream_t stream);
USL_API int usl_mc_emit(const usl_mc_args_t *a, usl_stream_t stream);
/* out[i] = sum_{j<i} f(in[j]), f = popcount or identity; total[0] = sum of all; block_sums: workspace of usl_scan_u8_blocks(n) uint32 */
USL_API int usl_scan_u8(const uint8_t *in, int64_t n, int popcount, uint32_t *out, uint32_t *block_sums, uint32_t *total,
                        usl_stream_t stream);
USL_API int usl_scan_u8_blocks(int64_t n, int64_t *n_blocks);

/* ---- f4: mesh culling after marching cubes (src/tools/cull_mesh.py:31-148; called from src/Mapper.py:556,570 and
 * src/utils/Mesher.py:274) ---------------------------------------------------------------------------------------------------
 * usl_mesh_cull_frames = the frame loop of cull_mesh (cull_mesh.py:58-99): seen[v] |= "some frame k has vertex v inside its
 * image (0 < u < W, 0 < v < H, edge = 0), in front of the camera (0 <= -z), and -- eval_rec, cfg['meshing']['eval_rec'] -- not
 * behind the bilinearly sampled sensor depth + truncation".  Projection as the reference writes it: cam = w2c @ [p,1], x negated,
 * uv = K @ cam, z = uv_z + 1e-5, depth sampled by F.grid_sample(align_corners=True, zeros) at 2*(u/W, v/H) - 1.
 * seen is OR-accumulated: the caller zero-fills it once and may call again per range of frames (whole_mask = !seen). */
typedef struct usl_cull_frames_args {
    const float *verts;               /* [V,3] mesh vertices as stored in the PLY (world frame, already divided by scale) */
    int64_t V;
    const float *w2c;                 /* [K,4,4] torch.inverse(c2w) of every frame (estimated or ground-truth poses, cull_mesh.py:63-69) */
    const float *depths;              /* [K,H,W] sensor depth frames; may be null when eval_rec == 0 */
    int32_t K, H, W;
    float fx, fy, cx, cy, truncation;
    int32_t eval_rec;
    int32_t frames_per_cta;           /* frames one CTA tests before it moves on (their depth images share L2): 0 = default 16, max 64 */
    uint8_t *seen;                    /* [V] in/out */
} usl_cull_frames_args_t;
USL_API int usl_mesh_cull_frames(const usl_cull_frames_args_t *a, usl_stream_t stream);
/* cull_out_bound_mesh's mesh_bound.contains(vertices) (cull_mesh.py:137-143) for a closed CONVEX bound (the reference's bound is
 * the convex hull Mesher.get_bound_from_frames returns): inside[v] = all_f (planes[f,0:3] . p + planes[f,3] <= 0); planes[F,4]
 * outward, on the device. */
USL_API int usl_mesh_cull_hull(const float *verts, int64_t V, const float *planes, int32_t F, uint8_t *inside, usl_stream_t stream);
/* The face rule + trimesh's update_faces / remove_unreferenced_vertices (cull_mesh.py:100-103, 144-146), order-preserving:
 *   usl_mesh_face_keep: keep[t] = any (require_all = 0, cull_mesh: vmask = seen) / all (require_all = 1, bound: vmask = inside) of
 *                       the face's three vertex flags; vref[v] = 1 for every vertex of a kept face (caller zero-fills vref)
 *   usl_scan_u8(keep) -> foff, T';  usl_scan_u8(vref) -> voff, V'   (popcount = 0)
 *   usl_mesh_compact:   verts_out[voff[v]] = verts[v] (+ colours, [V,3] u8, both or neither), faces_out[foff[t]] = voff[faces[t]] */
USL_API int usl_mesh_face_keep(const int32_t *faces, int64_t T, const uint8_t *vmask, int64_t V, int32_t require_all, uint8_t *keep,
                               uint8_t *vref, usl_stream_t stream);
USL_API int usl_mesh_compact(const float *verts, const uint8_t *colors, int64_t V, const int32_t *faces, int64_t T,
                             const uint8_t *keep, const uint8_t *vref, const uint32_t *voff, const uint32_t *foff,
                             float *verts_out, uint8_t *colors_out, int32_t *faces_out, usl_stream_t stream);

/* ---- f3: per-frame metrics of eval_rendering (src/tools/eval_recon.py:276-286) over the renderer's output buffers ----------
 * acc[0] += sum over the pixels with gt_depth > 0 of sum_c (gt_color - color)^2, acc[1] += sum |gt_depth - depth| over the same
 * pixels, acc[2] += their number -- in double (the reference's colour is float64 and render_img returns float64 depth).
 * mse = acc[0] / (3 acc[2]), psnr = -10 log10(mse), depth_l1 = acc[1] / acc[2].  gt_color / color [n,3], gt_depth / depth [n];
 * acc[3] is accumulated into (the caller zero-fills it; one slot of 3 doubles per frame keeps a sequence on the device). */
USL_API int usl_render_metrics(const float *gt_color, const float *gt_depth, const float *color, const float *depth, int64_t n,
                               double *acc, usl_stream_t stream);

/* ---- 8e: multi-GPU exchange steps of the sharded mapping iteration, over peer memory (NVLink / NVSwitch) -----------
 * The reference is single-GPU; these entry points are what its proposed `allreduce_grads` seam becomes.
 * The host layer maps every rank's buffers into every rank's address space (CUDA IPC / symmetric memory: one allocation
 * call per buffer at start-up) and passes the mapped pointers; nothing here calls NCCL. */
#define USL_MAX_PEERS 8
#define USL_PEER_CHANNELS 4
#define USL_PEER_CTRL_BYTES 2048          /* per-rank control block (flags, exchange slots, epochs): peer-mapped, zeroed once */
typedef struct usl_peers {
    int32_t rank, world;
    void *buf[USL_MAX_PEERS];             /* buf[p]: rank p's flat gradient buffer as mapped HERE (buf[rank] = the local one) */
    void *ctrl[USL_MAX_PEERS];            /* ctrl[p]: rank p's control block as mapped here */
    void *mc;                             /* multicast (NVLS) mapping of the same buffer, or NULL: when set, the reductions run as
                                           * multimem.ld_reduce (summed inside the NVSwitch) + multimem.st (replicated by the switch) */
    int32_t channel;                      /* barrier channel 0..USL_PEER_CHANNELS-1: exchanges issued concurrently on different
                                           * streams must use different channels (every rank the same one for the same exchange) */
    int32_t max_ctas_per_sm;              /* grid cap of the reduction kernels (0 = default 8); 1 leaves the SMs to a compute kernel
                                           * running beside the exchange (the pass is bound by the links, not by the SMs) */
} usl_peers_t;
USL_API int usl_peer_ctrl_bytes(void);
/* acc[USL_LOSS_SLOTS] (device, local) <- sum over ranks of acc: the loss sums / counts of usl_loss_fwd become global, so
 * every mean of the loss divides by the GLOBAL element count (one tiny kernel: push to all peers, flag, wait, sum) */
USL_API int usl_exchange_sums(const usl_peers_t *P, float *acc, usl_stream_t stream);
USL_API int usl_peer_barrier(const usl_peers_t *P, usl_stream_t stream);
/* buf[p][offset : offset+n] <- sum over ranks, on every rank (two-shot over peer memory: rank r reduces slice r and
 * pushes it to all; barrier before and after).  offset and n in floats, multiples of 4. */
USL_API int usl_allreduce_sum(const usl_peers_t *P, int64_t offset_floats, int64_t n_floats, usl_stream_t stream);
/* The same pass with the optimiser fused in (f1): the owner of a slice sums the gradients, applies torch.optim.Adam's update
 * (no weight decay / amsgrad) to its slice of the parameters and pushes the NEW PARAMETERS to every rank's param[p]
 * (same flat layout as the gradient buffers).  exp_avg / exp_avg_sq: this rank's state for its own slice only
 * (usl_allreduce_adam_slice_floats() floats each, zero-initialised).  ranges: learning rate per [begin,end) of the flat
 * layout (floats outside every range are left untouched).  param_mc: multicast mapping of the parameter buffers (or NULL).
 * step / step_dev as in usl_adam_step. */
#define USL_ADAM_MAX_RANGES 8
typedef struct usl_adam_range {
    int64_t begin, end;
    float lr, _pad;
} usl_adam_range_t;
USL_API int usl_allreduce_adam_slice_floats(int world, int64_t n_floats, int64_t *slice_floats);
USL_API int usl_allreduce_adam_step(const usl_peers_t *P, float *const *param, float *param_mc, int64_t offset_floats, int64_t n_floats,
                                    float *exp_avg, float *exp_avg_sq, const usl_adam_range_t *ranges, int n_ranges,
                                    float beta1, float beta2, float eps, int64_t step, const int64_t *step_dev,
                                    usl_stream_t stream);

/* ---- measurement utilities (no reference counterpart): ceilings for the roofline discussion ---- */
/* n_threads threads each issue per_thread random 8-byte loads from / vector atomics into table[entries*2] */
/* `repeats` coalesced read passes over buf[n_floats] (ld.global.cg, 16 bytes per lane): L2 read bandwidth when the buffer
 * fits L2, HBM read bandwidth when it does not */
USL_API int usl_bench_stream_read(const float *buf, int64_t n_floats, int repeats, float *out, usl_stream_t stream);
USL_API int usl_bench_gather(const float *table, uint32_t entries, int64_t n_threads, int per_thread, float *out,
                             usl_stream_t stream);
USL_API int usl_bench_scatter(float *table, uint32_t entries, int64_t n_threads, int per_thread, int mode,
                              usl_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* UNISLAM_B200_H */
