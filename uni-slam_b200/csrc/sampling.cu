// Ray generation, ray/bbox prefilter and z sampling.
// Replaces get_samples / get_samples_all / get_rays (src/common.py:95-180,210-228), the bbox prefilter
// (src/Mapper.py:396-401, src/Tracker.py:177-183), Renderer.perturbation + the depth-guided sampling
// (src/utils/Renderer.py:42-57,77-101) and the no-depth importance branch (Renderer.py:103-130 with
// common.sample_pdf, common.py:49-85).  RNG draws (torch.randint / torch.rand) are made by the host
// code exactly as in the reference and passed in as tensors.
#include "usl_field.cuh"

namespace usl {

// rays_d[a] = sum_b dir[b] * R[a][b]   (torch.sum(dirs[...,None,:] * c2w[:3,:3], -1), common.py:103,161)
// summed left to right like a sequential reduction over the length-3 axis.
__device__ __forceinline__ float rot_row(const float *__restrict__ c2w, int a, float d0, float d1, float d2) {
    const float p0 = __fmul_rn(d0, c2w[a * 4 + 0]);
    const float p1 = __fmul_rn(d1, c2w[a * 4 + 1]);
    const float p2 = __fmul_rn(d2, c2w[a * 4 + 2]);
    return __fadd_rn(__fadd_rn(p0, p1), p2);
}

__global__ void __launch_bounds__(256) sample_keyframe_rays_kernel(
    const float *__restrict__ c2ws, const float *__restrict__ depths, const float *__restrict__ colors,
    const float *__restrict__ dirs_cam, const int64_t *__restrict__ indices, int K, int64_t P, int n, int frame_base,
    float *__restrict__ rays_o, float *__restrict__ rays_d, float *__restrict__ gt_depth, float *__restrict__ gt_color,
    float *__restrict__ dirs_out, int32_t *__restrict__ frame_id) {
    const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= (int64_t)K * n) return;
    const int k = (int)(m / n);
    const int64_t src = (int64_t)k * P + indices[m];
    const float *c2w = c2ws + (int64_t)k * 16;
    const float d0 = dirs_cam[src * 3 + 0], d1 = dirs_cam[src * 3 + 1], d2 = dirs_cam[src * 3 + 2];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        rays_d[m * 3 + a] = rot_row(c2w, a, d0, d1, d2);
        rays_o[m * 3 + a] = c2w[a * 4 + 3];
        gt_color[m * 3 + a] = colors[src * 3 + a];
    }
    gt_depth[m] = depths[src];
    if (dirs_out) { dirs_out[m * 3 + 0] = d0; dirs_out[m * 3 + 1] = d1; dirs_out[m * 3 + 2] = d2; }
    if (frame_id) frame_id[m] = frame_base + k;
}

// camera-frame direction of pixel (i,j): ((i-cx)/fx, -(j-cy)/fy, -1)  (common.py:100)
__device__ __forceinline__ void pixel_dir(float i, float j, float fx, float fy, float cx, float cy, float d[3]) {
    d[0] = __fdiv_rn(__fsub_rn(i, cx), fx);
    d[1] = -__fdiv_rn(__fsub_rn(j, cy), fy);
    d[2] = -1.0f;
}

__global__ void __launch_bounds__(256) sample_window_rays_kernel(
    const float *__restrict__ c2w, const float *__restrict__ depth, const float *__restrict__ color, int H, int W,
    int H0, int H1, int W0, int W1, float fx, float fy, float cx, float cy, const int64_t *__restrict__ indices,
    int64_t n, float *__restrict__ rays_o, float *__restrict__ rays_d, float *__restrict__ gt_depth,
    float *__restrict__ gt_color, float *__restrict__ dirs_out) {
    const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= n) return;
    const int64_t idx = indices[m];
    const int Wc = W1 - W0;
    const int pi = (int)(idx % Wc) + W0, pj = (int)(idx / Wc) + H0;
    float d[3];
    pixel_dir((float)pi, (float)pj, fx, fy, cx, cy, d);
    const int64_t src = (int64_t)pj * W + pi;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        rays_d[m * 3 + a] = rot_row(c2w, a, d[0], d[1], d[2]);
        rays_o[m * 3 + a] = c2w[a * 4 + 3];
        gt_color[m * 3 + a] = color[src * 3 + a];
        if (dirs_out) dirs_out[m * 3 + a] = d[a];
    }
    gt_depth[m] = depth[src];
}

__global__ void __launch_bounds__(256) image_rays_kernel(const float *__restrict__ c2w, int H, int W, float fx, float fy,
                                                         float cx, float cy, float *__restrict__ rays_o,
                                                         float *__restrict__ rays_d) {
    const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= (int64_t)H * W) return;
    float d[3];
    pixel_dir((float)(m % W), (float)(m / W), fx, fy, cx, cy, d);
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        rays_d[m * 3 + a] = rot_row(c2w, a, d[0], d[1], d[2]);
        rays_o[m * 3 + a] = c2w[a * 4 + 3];
    }
}

// t_exit = min_d max((lo_d - o_d)/d_d, (hi_d - o_d)/d_d)
__device__ __forceinline__ float bbox_exit(const float o[3], const float d[3], const usl_bound_t &b) {
    float t = INFINITY;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const float t0 = __fdiv_rn(__fsub_rn(b.lo[a], o[a]), d[a]);
        const float t1 = __fdiv_rn(__fsub_rn(b.hi[a], o[a]), d[a]);
        // torch.max / torch.min propagate NaN (0/0 when a ray is parallel to and on a face)
        float m = fmaxf(t0, t1);
        if (t0 != t0 || t1 != t1) m = NAN;
        t = (m != m || t != t) ? NAN : fminf(t, m);
    }
    return t;
}

__global__ void __launch_bounds__(256) bbox_prefilter_kernel(const float *__restrict__ rays_o, const float *__restrict__ rays_d,
                                                             const float *__restrict__ gt_depth, int64_t n, usl_bound_t b,
                                                             int require_depth, float *__restrict__ t_exit,
                                                             uint8_t *__restrict__ valid) {
    const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= n) return;
    const float o[3] = {rays_o[m * 3], rays_o[m * 3 + 1], rays_o[m * 3 + 2]};
    const float d[3] = {rays_d[m * 3], rays_d[m * 3 + 1], rays_d[m * 3 + 2]};
    const float t = bbox_exit(o, d, b);
    if (t_exit) t_exit[m] = t;
    if (valid) {
        const float g = gt_depth[m];
        bool v = t >= g;
        if (require_depth) v = v && (g > 0.f);
        valid[m] = v ? 1 : 0;
    }
}

// ---- depth-guided z sampling: one thread per (ray, sample); merge by rank, perturb through shared memory ----
#define ZS_RAYS_PER_BLOCK 8
__global__ void zsample_depth_kernel(usl_zsample_args_t a, const float *__restrict__ gt_depth,
                                     const uint8_t *__restrict__ valid, const float *__restrict__ t_rand,
                                     const int32_t *__restrict__ row_map, int64_t n_rays, float *__restrict__ z) {
    extern __shared__ float sz[];   // [ZS_RAYS_PER_BLOCK][S] merged (sorted) samples
    const int ns = a.n_stratified, ni = a.n_importance, S = ns + ni;
    const int rl = threadIdx.x / S, k = threadIdx.x - rl * S;
    const int64_t r = (int64_t)blockIdx.x * ZS_RAYS_PER_BLOCK + rl;
    float gt = 0.f;
    bool live = r < n_rays && (!valid || valid[r]);
    if (live) { gt = gt_depth[r]; live = gt > 0.f; }
    // z_free_i = 0.0 + (1.2*gt)*t_uni[i] ; z_surf_j = (gt - 1.5tr) + (3tr * t_surf[j])     (Renderer.py:91-95)
    const float g12 = __fmul_rn(1.2f, gt);
    const float s0 = __fsub_rn(gt, a.c_surf_lo);
    if (live) {
        float v;
        int rank;
        if (k < ns) {           // rank among the merged list: equal values keep the free sample first (== stable two-pointer merge)
            v = __fadd_rn(0.0f, __fmul_rn(g12, a.t_uni[k]));
            rank = k;
            for (int j = 0; j < ni; ++j) rank += (__fadd_rn(s0, __fmul_rn(a.c_surf_span, a.t_surf[j])) < v) ? 1 : 0;
        } else {
            const int j0 = k - ns;
            v = __fadd_rn(s0, __fmul_rn(a.c_surf_span, a.t_surf[j0]));
            rank = j0;
            for (int i = 0; i < ns; ++i) rank += (__fadd_rn(0.0f, __fmul_rn(g12, a.t_uni[i])) <= v) ? 1 : 0;
        }
        sz[rl * S + rank] = v;
    }
    __syncthreads();
    if (!live) return;
    const float *zr = sz + rl * S;
    float out = zr[k];
    if (t_rand) {   // Renderer.perturbation (Renderer.py:42-57)
        const float cur = out;
        const float lower = (k == 0) ? cur : __fmul_rn(0.5f, __fadd_rn(cur, zr[k - 1]));
        const float upper = (k == S - 1) ? cur : __fmul_rn(0.5f, __fadd_rn(zr[k + 1], cur));
        const float t = t_rand[(int64_t)(row_map ? row_map[r] : r) * S + k];
        out = __fadd_rn(lower, __fmul_rn(__fsub_rn(upper, lower), t));
    }
    z[r * S + k] = out;
}

// ---- fused ray set-up: ray generation + bbox prefilter + depth-guided z sampling in one launch ----
// Block = ZS_RAYS_PER_BLOCK rays x S threads; thread k == 0 of each ray gathers / rotates the ray and tests the bbox,
// then all S threads of the ray produce its z samples.  With cam_poses the camera matrix is rebuilt from the pose in
// place (same arithmetic as usl_pose_to_matrix).
__global__ void ray_setup_kernel(const __grid_constant__ usl_ray_setup_t A) {
    extern __shared__ float sz[];                      // [ZS_RAYS_PER_BLOCK][S] merged samples
    __shared__ float s_gt[ZS_RAYS_PER_BLOCK];
    __shared__ int s_live[ZS_RAYS_PER_BLOCK];
    const int ns = A.zs.n_stratified, ni = A.zs.n_importance, S = ns + ni;
    const int rl = threadIdx.x / S, k = threadIdx.x - rl * S;
    const int64_t r = (int64_t)blockIdx.x * ZS_RAYS_PER_BLOCK + rl;
    if (k == 0) {
        int live = 0;
        float gt = 0.f;
        if (r < A.n_rays) {
            float c2w[12], d[3], col[3];
            int frame = 0;
            if (A.mode == 0) {
                int b = 0;
                int64_t m = r + A.ray_offset;                        // global ray slot (a rank of a sharded batch starts mid-list)
                const int64_t n0 = (int64_t)A.batch[0].K * A.batch[0].n;
                if (A.n_batches > 1 && m >= n0) { b = 1; m -= n0; }
                const usl_ray_batch_t &B = A.batch[b];
                const int kf = (int)(m / B.n);
                frame = B.frame_base + kf;
                const int64_t src = (int64_t)kf * B.P + B.indices[m];
                d[0] = B.dirs_cam[src * 3]; d[1] = B.dirs_cam[src * 3 + 1]; d[2] = B.dirs_cam[src * 3 + 2];
                col[0] = B.colors[src * 3]; col[1] = B.colors[src * 3 + 1]; col[2] = B.colors[src * 3 + 2];
                gt = B.depths[src];
                if (A.cam_poses) {
                    if (frame == 0) {
#pragma unroll
                        for (int q = 0; q < 12; ++q) c2w[q] = A.c2w_fixed[q];
                    } else pose_to_c2w(A.cam_poses + (int64_t)(frame - 1) * 7, c2w);
                } else {
#pragma unroll
                    for (int q = 0; q < 12; ++q) c2w[q] = B.c2ws[(int64_t)kf * 16 + q];
                }
            } else {
                int pi, pj;
                if (A.mode == 2) {                                   // whole frame, row-major (get_rays, common.py:210-228)
                    const int64_t pix = A.pixel_begin + r;
                    pi = (int)(pix % A.W); pj = (int)(pix / A.W);
                } else {
                    const int64_t idx = A.win_indices[r];
                    const int Wc = A.W1 - A.W0;
                    pi = (int)(idx % Wc) + A.W0; pj = (int)(idx / Wc) + A.H0;
                }
                pixel_dir((float)pi, (float)pj, A.fx, A.fy, A.cx, A.cy, d);
                const int64_t src = (int64_t)pj * A.W + pi;
                if (A.color_img) { col[0] = A.color_img[src * 3]; col[1] = A.color_img[src * 3 + 1]; col[2] = A.color_img[src * 3 + 2]; }
                else col[0] = col[1] = col[2] = 0.f;
                gt = A.depth_img[src];
                if (A.cam_poses) pose_to_c2w(A.cam_poses, c2w);
                else {
#pragma unroll
                    for (int q = 0; q < 12; ++q) c2w[q] = A.c2w[q];
                }
            }
            float o[3], rd[3];
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                rd[a] = __fadd_rn(__fadd_rn(__fmul_rn(d[0], c2w[a * 4]), __fmul_rn(d[1], c2w[a * 4 + 1])), __fmul_rn(d[2], c2w[a * 4 + 2]));
                o[a] = c2w[a * 4 + 3];
                A.rays_d[r * 3 + a] = rd[a]; A.rays_o[r * 3 + a] = o[a];
                if (A.gt_color) A.gt_color[r * 3 + a] = col[a];
                if (A.dirs_out) A.dirs_out[r * 3 + a] = d[a];
            }
            A.gt_depth[r] = gt;
            if (A.frame_id) A.frame_id[r] = frame;
            bool v = (A.mode == 2) || bbox_exit(o, rd, A.bound) >= gt;
            if (A.require_depth) v = v && (gt > 0.f);
            A.valid[r] = v ? 1 : 0;
            live = v ? ((gt > 0.f) ? 1 : 2) : 0;           // 2: valid ray without sensor depth (usl_zsample_nodepth fills its z row)
        }
        s_gt[rl] = gt; s_live[rl] = live;
    }
    __syncthreads();
    const bool live = s_live[rl] == 1;
    const float gt = s_gt[rl];
    const float g12 = __fmul_rn(1.2f, gt);
    const float s0 = __fsub_rn(gt, A.zs.c_surf_lo);
    if (live) {
        float v;
        int rank;
        if (k < ns) {
            v = __fadd_rn(0.0f, __fmul_rn(g12, A.zs.t_uni[k]));
            rank = k;
            for (int j = 0; j < ni; ++j) rank += (__fadd_rn(s0, __fmul_rn(A.zs.c_surf_span, A.zs.t_surf[j])) < v) ? 1 : 0;
        } else {
            const int j0 = k - ns;
            v = __fadd_rn(s0, __fmul_rn(A.zs.c_surf_span, A.zs.t_surf[j0]));
            rank = j0;
            for (int i = 0; i < ns; ++i) rank += (__fadd_rn(0.0f, __fmul_rn(g12, A.zs.t_uni[i])) <= v) ? 1 : 0;
        }
        sz[rl * S + rank] = v;
    }
    __syncthreads();
    if (!live) {
        // a depth-less valid ray must not keep a z row from an earlier iteration: poison it, so a caller that skips
        // usl_zsample_nodepth sees NaN instead of silently stale samples
        if (s_live[rl] == 2) A.z[r * S + k] = __int_as_float(0x7fc00000);
        return;
    }
    const float *zr = sz + rl * S;
    float out = zr[k];
    if (A.t_rand) {
        const float cur = out;
        const float lower = (k == 0) ? cur : __fmul_rn(0.5f, __fadd_rn(cur, zr[k - 1]));
        const float upper = (k == S - 1) ? cur : __fmul_rn(0.5f, __fadd_rn(zr[k + 1], cur));
        out = __fadd_rn(lower, __fmul_rn(__fsub_rn(upper, lower), A.t_rand[(r + A.ray_offset) * S + k]));
    }
    A.z[r * S + k] = out;
}

// ---- no-depth rays: uniform samples to the bbox exit + inverse-CDF resampling from an SDF query ----
#define ND_MAX_STRAT 64
#define ND_MAX_IMP 32
struct NoDepthArgs {
    usl_zsample_args_t a;
    usl_field_t f;
    const float *beta, *rays_o, *rays_d, *gt_depth, *t_rand_uni, *u_pdf;
    const uint8_t *valid;
    const int32_t *row_map;
    int64_t n_rays;
    float *z;
    int64_t *pdf_inds;
};

// SDF-only decode for latency-bound callers (a few hundred rays): the gathers of B levels are issued back to back, so a
// point pays 16/B memory round trips instead of 16.  Same arithmetic, in the same order, as decode_point<false,false>
// (nested-lerp interpolation, first layer accumulated level by level), hence the same values.
template <int B>
__device__ __forceinline__ float decode_sdf_batched(const usl_grid_t &g, const float2 *__restrict__ table, const usl_mlp_t &m,
                                                    const MlpSmem &sm, const float xc[3]) {
    float h[USL_HID];
#pragma unroll
    for (int j = 0; j < USL_HID; ++j) h[j] = sm.b1[j];
    for (int l0 = 0; l0 + B <= g.n_levels; l0 += B) {
        float2 v[B][8];
        float w[B][3];
#pragma unroll
        for (int q = 0; q < B; ++q) {
            const usl_level_t &lv = g.levels[l0 + q];
            const Cell c = make_cell(lv, xc[0], xc[1], xc[2]);
            uint32_t idx[8];
            corner_indices<true>(lv, c, idx);
            const float2 *tab = table + lv.offset;
#pragma unroll
            for (int k = 0; k < 8; ++k) v[q][k] = ldg2(tab + idx[k]);
            w[q][0] = c.w[0]; w[q][1] = c.w[1]; w[q][2] = c.w[2];
        }
#pragma unroll
        for (int q = 0; q < B; ++q) {
            float2 a[4], b[2];
#pragma unroll
            for (int p = 0; p < 4; ++p)
                a[p] = make_float2(fmaf(w[q][0], v[q][2 * p + 1].x - v[q][2 * p].x, v[q][2 * p].x),
                                   fmaf(w[q][0], v[q][2 * p + 1].y - v[q][2 * p].y, v[q][2 * p].y));
#pragma unroll
            for (int z = 0; z < 2; ++z)
                b[z] = make_float2(fmaf(w[q][1], a[2 * z + 1].x - a[2 * z].x, a[2 * z].x), fmaf(w[q][1], a[2 * z + 1].y - a[2 * z].y, a[2 * z].y));
            const float2 f = make_float2(fmaf(w[q][2], b[1].x - b[0].x, b[0].x), fmaf(w[q][2], b[1].y - b[0].y, b[0].y));
            const int l = l0 + q;
            const float4 *wa = reinterpret_cast<const float4 *>(sm.w1t[2 * l]);
            const float4 *wb = reinterpret_cast<const float4 *>(sm.w1t[2 * l + 1]);
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const float4 a4 = wa[r], b4 = wb[r];
                const float av[4] = {a4.x, a4.y, a4.z, a4.w}, bv[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int j = r * 4 + e;
                    h[j] = fmaf(av[e], f.x, h[j]);
                    h[j] = fmaf(bv[e], f.y, h[j]);
                }
            }
        }
    }
    float th[1][USL_HID], out[4], tout[4][3];
    mlp_tail<false>(m, sm, h, th, out, tout);
    return out[0];
}

// ND_WARPS rays per CTA.  Depth-less rays are a few per cent of a batch, scattered, and each is one long dependent chain
// (ray -> stratified z -> two rounds of 8 levels of gathers -> decoder -> sequential cdf -> searchsorted -> rank merge), so the
// kernel's duration is that chain's latency: 1, 2 and 4 rays per CTA all measure 25 us on the mapping workload.
#ifndef ND_WARPS
#define ND_WARPS 4
#endif
__global__ void __launch_bounds__(ND_WARPS * 32) zsample_nodepth_kernel(const __grid_constant__ NoDepthArgs A) {
    __shared__ MlpSmem sm;
    __shared__ float s_z[ND_WARPS][ND_MAX_STRAT], s_w[ND_WARPS][ND_MAX_STRAT], s_cdf[ND_WARPS][ND_MAX_STRAT], s_smp[ND_WARPS][ND_MAX_IMP];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t r = (int64_t)blockIdx.x * ND_WARPS + warp;
    // only rays with gt_depth == 0 take this branch (gt_mask, Renderer.py:83,104): skip the weight staging otherwise
    const bool mine = r < A.n_rays && (!A.valid || A.valid[r]) && !(A.gt_depth[r] > 0.f);
    if (!__syncthreads_or(mine ? 1 : 0)) return;
    stage_mlp(A.f.mlp[0], sm);
    __syncthreads();
    if (!mine) return;
    const int ns = A.a.n_stratified, ni = A.a.n_importance, S = ns + ni;
    const int row = A.row_map ? A.row_map[r] : (int)r;
    float *zs = s_z[warp], *ws = s_w[warp], *cdf = s_cdf[warp], *smp = s_smp[warp];
    const float o[3] = {A.rays_o[r * 3], A.rays_o[r * 3 + 1], A.rays_o[r * 3 + 2]};
    const float d[3] = {A.rays_d[r * 3], A.rays_d[r * 3 + 1], A.rays_d[r * 3 + 2]};
    usl_bound_t b;
#pragma unroll
    for (int q = 0; q < 3; ++q) { b.lo[q] = A.f.bound_lo[q]; b.hi[q] = A.f.bound_hi[q]; }
    const float far_bb = __fadd_rn(bbox_exit(o, d, b), 0.01f);                       // Renderer.py:110-113
    const float beta = A.beta[0];
    // z_uni = near*(1-t) + far*t with near = 0.0                                      (Renderer.py:115)
    for (int k = lane; k < ns; k += 32) {
        const float t = A.a.t_uni[k];
        zs[k] = __fadd_rn(__fmul_rn(0.0f, __fsub_rn(1.0f, t)), __fmul_rn(far_bb, t));
    }
    __syncwarp();
    if (A.t_rand_uni) {
        float pert[(ND_MAX_STRAT + 31) / 32];
        int q = 0;
        for (int k = lane; k < ns; k += 32, ++q) {
            const float cur = zs[k];
            const float lower = (k == 0) ? cur : __fmul_rn(0.5f, __fadd_rn(cur, zs[k - 1]));
            const float upper = (k == ns - 1) ? cur : __fmul_rn(0.5f, __fadd_rn(zs[k + 1], cur));
            pert[q] = __fadd_rn(lower, __fmul_rn(__fsub_rn(upper, lower), A.t_rand_uni[(int64_t)row * ns + k]));
        }
        __syncwarp();
        q = 0;
        for (int k = lane; k < ns; k += 32, ++q) zs[k] = pert[q];
        __syncwarp();
    }
    // SDF query at coords normalised to [-1,1] then clamped to [0,1] (Renderer.py:120, common.py:231-245, decoders.py:101)
    for (int k = lane; k < ns; k += 32) {
        float xc[3];
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            const float p = __fadd_rn(o[q], __fmul_rn(d[q], zs[k]));
            const float xn = __fsub_rn(__fmul_rn(__fdiv_rn(__fsub_rn(p, b.lo[q]), __fsub_rn(b.hi[q], b.lo[q])), 2.0f), 1.0f);
            xc[q] = fminf(fmaxf(xn, 0.f), 1.f);
        }
        const float sdf = decode_sdf_batched<8>(A.f.grid[0], reinterpret_cast<const float2 *>(A.f.table[0]), A.f.mlp[0], sm, xc);
        const float sg = 1.0f / (1.0f + expf(sdf * beta));                           // sigmoid(-sdf*beta)
        ws[k] = 1.0f - expf(-beta * sg);                                             // alpha (Renderer.py:154-158)
    }
    __syncwarp();
    // weights = alpha * exclusive cumprod(1 - alpha + 1e-10); cdf = [0, cumsum(weights[1:-1])] -- sequential like torch
    if (lane == 0) {
        float T = 1.0f;
        for (int k = 0; k < ns; ++k) {
            const float al = ws[k];
            ws[k] = __fmul_rn(al, T);
            T = __fmul_rn(T, __fadd_rn(__fsub_rn(1.0f, al), 1e-10f));
        }
        float c = 0.f;
        cdf[0] = 0.f;
        for (int k = 1; k <= ns - 2; ++k) { c = __fadd_rn(c, ws[k]); cdf[k] = c; }
    }
    __syncwarp();
    const int ncdf = ns - 1;                                  // == number of bins (z_mid)
    for (int q = lane; q < ni; q += 32) {
        const float u = A.u_pdf[(int64_t)row * ni + q];
        int lo = 0, hi = ncdf;                               // searchsorted(cdf, u, right=True): #elements <= u
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (cdf[mid] <= u) lo = mid + 1; else hi = mid; }
        const int inds = lo;
        if (A.pdf_inds) A.pdf_inds[(int64_t)row * ni + q] = inds;
        const int below = max(inds - 1, 0), above = min(ncdf - 1, inds);
        float denom = __fsub_rn(cdf[above], cdf[below]);
        if (denom < 1e-5f) denom = 1.0f;
        const float t = __fdiv_rn(__fsub_rn(u, cdf[below]), denom);
        const float bb = __fmul_rn(0.5f, __fadd_rn(zs[below + 1], zs[below]));
        const float ba = __fmul_rn(0.5f, __fadd_rn(zs[above + 1], zs[above]));
        smp[q] = __fadd_rn(bb, __fmul_rn(t, __fsub_rn(ba, bb)));
    }
    __syncwarp();
    // sort(cat(z_uni, samples)): rank every element among both lists (ties: z_uni first, then by index)
    float *zr = A.z + r * S;
    for (int k = lane; k < S; k += 32) {
        const bool from_uni = k < ns;
        const float v = from_uni ? zs[k] : smp[k - ns];
        int rank = 0;
        for (int q = 0; q < ns; ++q) {
            const float w = zs[q];
            rank += (w < v) || (w == v && (!from_uni || q < k));
        }
        for (int q = 0; q < ni; ++q) {
            const float w = smp[q];
            rank += (w < v) || (w == v && !from_uni && q < k - ns);
        }
        zr[rank] = v;
    }
}

// ---- f2: keyframe store (src/Mapper.py:528-541) and co-visibility (Mapper.py:177-236) --------------------------------
// A frame becomes a keyframe as a 10 % pixel subset: one launch gathers colour / depth / camera direction of the drawn pixels
// into the store's row (the reference does three boolean-free index ops + a dict of tensors per keyframe).
__global__ void __launch_bounds__(256) keyframe_insert_kernel(const float *__restrict__ color, const float *__restrict__ depth,
                                                              const float *__restrict__ dirs, const int64_t *__restrict__ indices,
                                                              int64_t P, float *__restrict__ color_row, float *__restrict__ depth_row,
                                                              float *__restrict__ dirs_row, int64_t *__restrict__ idx_row) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    const int64_t src = indices[i];
    depth_row[i] = depth[src];
#pragma unroll
    for (int a = 0; a < 3; ++a) { color_row[i * 3 + a] = color[src * 3 + a]; dirs_row[i * 3 + a] = dirs[src * 3 + a]; }
    if (idx_row) idx_row[i] = src;
}

// keyframe_selection_LC's overlap measure: sample num_samples points in [0.8 d, d + 0.5] along every ray with sensor depth,
// project them into every keyframe and report the fraction that lands inside the image (minus an edge) in front of the camera.
// One CTA per keyframe; the world-to-camera matrix is the closed-form inverse of the rigid c2w (R^T, -R^T t) in double.
struct CovisArgs {
    const float *rays_o, *rays_d, *gt_depth;
    int64_t n;
    int num_samples;
    const float *c2ws;
    int K, H, W;
    float fx, fy, cx, cy, edge;
    float *percent_inside;
};
__global__ void __launch_bounds__(256) keyframe_covis_kernel(const __grid_constant__ CovisArgs A) {
    __shared__ float s_w2c[12];
    __shared__ int s_cnt[2];
    const int k = blockIdx.x;
    if (threadIdx.x == 0) {
        const float *c = A.c2ws + (int64_t)k * 16;
        // general 3x3 inverse (adjugate / determinant) in double: equals R^T for a rotation, and stays exact enough for the
        // un-normalised rotations the optimiser produces
        double m[9] = {c[0], c[1], c[2], c[4], c[5], c[6], c[8], c[9], c[10]};
        const double det = m[0] * (m[4] * m[8] - m[5] * m[7]) - m[1] * (m[3] * m[8] - m[5] * m[6]) + m[2] * (m[3] * m[7] - m[4] * m[6]);
        const double id = 1.0 / det;
        double inv[9] = {(m[4] * m[8] - m[5] * m[7]) * id, (m[2] * m[7] - m[1] * m[8]) * id, (m[1] * m[5] - m[2] * m[4]) * id,
                         (m[5] * m[6] - m[3] * m[8]) * id, (m[0] * m[8] - m[2] * m[6]) * id, (m[2] * m[3] - m[0] * m[5]) * id,
                         (m[3] * m[7] - m[4] * m[6]) * id, (m[1] * m[6] - m[0] * m[7]) * id, (m[0] * m[4] - m[1] * m[3]) * id};
        const double t[3] = {c[3], c[7], c[11]};
        for (int r = 0; r < 3; ++r) {
            s_w2c[r * 4 + 0] = (float)inv[r * 3]; s_w2c[r * 4 + 1] = (float)inv[r * 3 + 1]; s_w2c[r * 4 + 2] = (float)inv[r * 3 + 2];
            s_w2c[r * 4 + 3] = (float)(-(inv[r * 3] * t[0] + inv[r * 3 + 1] * t[1] + inv[r * 3 + 2] * t[2]));
        }
        s_cnt[0] = 0; s_cnt[1] = 0;
    }
    __syncthreads();
    int inside = 0, total = 0;
    const int64_t M = A.n * A.num_samples;
    for (int64_t j = threadIdx.x; j < M; j += blockDim.x) {
        const int64_t r = j / A.num_samples;
        const int q = (int)(j - r * A.num_samples);
        const float d = A.gt_depth[r];
        if (!(d > 0.f)) continue;                                        // nonzero_depth filter (Mapper.py:201-204)
        ++total;
        const float tv = (A.num_samples > 1) ? (float)q / (float)(A.num_samples - 1) : 0.f;     // torch.linspace(0, 1, num_samples)
        const float z = (d * 0.8f) * (1.0f - tv) + (d + 0.5f) * tv;
        float p[3];
#pragma unroll
        for (int a = 0; a < 3; ++a) p[a] = A.rays_o[r * 3 + a] + A.rays_d[r * 3 + a] * z;
        float cc[3];
#pragma unroll
        for (int a = 0; a < 3; ++a) cc[a] = s_w2c[a * 4] * p[0] + s_w2c[a * 4 + 1] * p[1] + s_w2c[a * 4 + 2] * p[2] + s_w2c[a * 4 + 3];
        cc[0] = -cc[0];                                                  // cam_cords[:, :, 0] *= -1
        const float zz = cc[2] + 1e-5f;                                  // uv = K @ cam; z = uv_z + 1e-5
        const float u = (A.fx * cc[0] + A.cx * cc[2]) / zz, v = (A.fy * cc[1] + A.cy * cc[2]) / zz;
        const bool in = (u < A.W - A.edge) && (u > A.edge) && (v < A.H - A.edge) && (v > A.edge) && (zz < 0.f);
        inside += in ? 1 : 0;
    }
    atomicAdd(&s_cnt[0], inside); atomicAdd(&s_cnt[1], total);
    __syncthreads();
    if (threadIdx.x == 0) A.percent_inside[k] = s_cnt[1] > 0 ? (float)s_cnt[0] / (float)s_cnt[1] : 0.f;
}

}  // namespace usl

using namespace usl;

extern "C" {

int usl_keyframe_insert(const float *color_img, const float *depth_img, const float *dirs_cam, const int64_t *indices, int64_t P,
                        float *color_row, float *depth_row, float *dirs_row, int64_t *idx_row, usl_stream_t stream) {
    if (P <= 0) return 0;
    if (!color_img || !depth_img || !dirs_cam || !indices || !color_row || !depth_row || !dirs_row) { set_error("usl_keyframe_insert: null argument"); return 1; }
    keyframe_insert_kernel<<<(unsigned)((P + 255) / 256), 256, 0, (cudaStream_t)stream>>>(color_img, depth_img, dirs_cam, indices, P, color_row,
                                                                                           depth_row, dirs_row, idx_row);
    return check_launch("usl_keyframe_insert");
}

int usl_keyframe_covisibility(const float *rays_o, const float *rays_d, const float *gt_depth, int64_t n, int num_samples,
                              const float *c2ws, int K, int H, int W, float fx, float fy, float cx, float cy, float edge,
                              float *percent_inside, usl_stream_t stream) {
    if (K <= 0) return 0;
    if (!rays_o || !rays_d || !gt_depth || !c2ws || !percent_inside || n < 0 || num_samples < 1) { set_error("usl_keyframe_covisibility: bad arguments"); return 1; }
    CovisArgs A{rays_o, rays_d, gt_depth, n, num_samples, c2ws, K, H, W, fx, fy, cx, cy, edge, percent_inside};
    keyframe_covis_kernel<<<(unsigned)K, 256, 0, (cudaStream_t)stream>>>(A);
    return check_launch("usl_keyframe_covisibility");
}

int usl_sample_keyframe_rays(const float *c2ws, const float *depths, const float *colors, const float *dirs_cam,
                             const int64_t *indices, int K, int64_t P, int n, int frame_base, float *rays_o,
                             float *rays_d, float *gt_depth, float *gt_color, float *dirs_out, int32_t *frame_id,
                             usl_stream_t stream) {
    const int64_t M = (int64_t)K * n;
    if (M <= 0) return 0;
    sample_keyframe_rays_kernel<<<(unsigned)((M + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        c2ws, depths, colors, dirs_cam, indices, K, P, n, frame_base, rays_o, rays_d, gt_depth, gt_color, dirs_out, frame_id);
    return check_launch("usl_sample_keyframe_rays");
}

int usl_sample_window_rays(const float *c2w, const float *depth, const float *color, int H, int W, int H0, int H1,
                           int W0, int W1, float fx, float fy, float cx, float cy, const int64_t *indices, int64_t n,
                           float *rays_o, float *rays_d, float *gt_depth, float *gt_color, float *dirs_out,
                           usl_stream_t stream) {
    if (n <= 0) return 0;
    if (H0 < 0 || H1 > H || W0 < 0 || W1 > W || H1 <= H0 || W1 <= W0) { set_error("usl_sample_window_rays: bad window"); return 1; }
    sample_window_rays_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        c2w, depth, color, H, W, H0, H1, W0, W1, fx, fy, cx, cy, indices, n, rays_o, rays_d, gt_depth, gt_color, dirs_out);
    return check_launch("usl_sample_window_rays");
}

int usl_image_rays(const float *c2w, int H, int W, float fx, float fy, float cx, float cy, float *rays_o,
                   float *rays_d, usl_stream_t stream) {
    const int64_t M = (int64_t)H * W;
    if (M <= 0) return 0;
    image_rays_kernel<<<(unsigned)((M + 255) / 256), 256, 0, (cudaStream_t)stream>>>(c2w, H, W, fx, fy, cx, cy, rays_o, rays_d);
    return check_launch("usl_image_rays");
}

int usl_bbox_prefilter(const float *rays_o, const float *rays_d, const float *gt_depth, int64_t n,
                       const usl_bound_t *bound, int require_depth, float *t_exit, uint8_t *valid,
                       usl_stream_t stream) {
    if (n <= 0) return 0;
    if (!bound || (valid && !gt_depth)) { set_error("usl_bbox_prefilter: bad arguments"); return 1; }
    bbox_prefilter_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(rays_o, rays_d, gt_depth, n, *bound,
                                                                                         require_depth, t_exit, valid);
    return check_launch("usl_bbox_prefilter");
}

int usl_ray_setup(const usl_ray_setup_t *a, usl_stream_t stream) {
    if (!a || a->n_rays <= 0) return a ? 0 : 1;
    const int S = a->zs.n_stratified + a->zs.n_importance;
    if (S < 2 || S > 128) { set_error("usl_ray_setup: n_stratified + n_importance must be in 2..128"); return 1; }
    if (a->mode == 0 && (a->n_batches < 1 || a->n_batches > 2)) { set_error("usl_ray_setup: 1 or 2 keyframe batches"); return 1; }
    if (a->ray_offset != 0 && a->mode != 0) { set_error("usl_ray_setup: ray_offset is a mode-0 (keyframe batches) feature"); return 1; }
    if (a->mode == 0) {
        int64_t total = 0;
        for (int b = 0; b < a->n_batches; ++b) total += (int64_t)a->batch[b].K * a->batch[b].n;
        if (a->ray_offset < 0 || a->ray_offset + a->n_rays > total) { set_error("usl_ray_setup: ray range outside the batches"); return 1; }
    }
    if (a->mode == 1 && (a->H0 < 0 || a->H1 > a->H || a->W0 < 0 || a->W1 > a->W || a->H1 <= a->H0 || a->W1 <= a->W0)) { set_error("usl_ray_setup: bad window"); return 1; }
    if (a->mode == 2 && (a->pixel_begin < 0 || a->pixel_begin + a->n_rays > (int64_t)a->H * a->W || !a->depth_img)) { set_error("usl_ray_setup: pixel range outside the frame"); return 1; }
    if (a->mode != 2 && (!a->gt_color || !a->dirs_out)) { set_error("usl_ray_setup: gt_color / dirs_out are required in modes 0 and 1"); return 1; }
    ray_setup_kernel<<<(unsigned)((a->n_rays + ZS_RAYS_PER_BLOCK - 1) / ZS_RAYS_PER_BLOCK), ZS_RAYS_PER_BLOCK * S,
                       ZS_RAYS_PER_BLOCK * S * sizeof(float), (cudaStream_t)stream>>>(*a);
    return check_launch("usl_ray_setup");
}

int usl_zsample_depth(const usl_zsample_args_t *a, const float *gt_depth, const uint8_t *valid, const float *t_rand,
                      const int32_t *row_map, int64_t n_rays, float *z, usl_stream_t stream) {
    if (n_rays <= 0) return 0;
    if (!a || a->n_stratified < 1 || a->n_importance < 1) { set_error("usl_zsample_depth: bad arguments"); return 1; }
    const int S = a->n_stratified + a->n_importance;
    if (S > 128) { set_error("usl_zsample_depth: n_stratified + n_importance must be <= 128"); return 1; }
    zsample_depth_kernel<<<(unsigned)((n_rays + ZS_RAYS_PER_BLOCK - 1) / ZS_RAYS_PER_BLOCK), ZS_RAYS_PER_BLOCK * S,
                           ZS_RAYS_PER_BLOCK * S * sizeof(float), (cudaStream_t)stream>>>(*a, gt_depth, valid, t_rand, row_map, n_rays, z);
    return check_launch("usl_zsample_depth");
}

int usl_zsample_nodepth(const usl_zsample_args_t *a, const usl_field_t *f, const float *beta, const float *rays_o,
                        const float *rays_d, const float *gt_depth, const uint8_t *valid, const float *t_rand_uni,
                        const float *u_pdf, const int32_t *row_map, int64_t n_rays, float *z, int64_t *pdf_inds,
                        usl_stream_t stream) {
    if (n_rays <= 0) return 0;
    if (!a || !f || a->n_stratified < 3 || a->n_stratified > ND_MAX_STRAT || a->n_importance < 1 || a->n_importance > ND_MAX_IMP || !u_pdf) {
        set_error("usl_zsample_nodepth: unsupported sample counts (n_stratified 3..64, n_importance 1..32)");
        return 1;
    }
    NoDepthArgs A;
    A.a = *a; A.f = *f; A.beta = beta; A.rays_o = rays_o; A.rays_d = rays_d; A.gt_depth = gt_depth;
    A.t_rand_uni = t_rand_uni; A.u_pdf = u_pdf; A.valid = valid; A.row_map = row_map; A.n_rays = n_rays; A.z = z; A.pdf_inds = pdf_inds;
    zsample_nodepth_kernel<<<(unsigned)((n_rays + ND_WARPS - 1) / ND_WARPS), ND_WARPS * 32, 0, (cudaStream_t)stream>>>(A);
    return check_launch("usl_zsample_nodepth");
}

}  // extern "C"
