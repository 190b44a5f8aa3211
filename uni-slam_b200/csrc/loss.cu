// Masks + uncertainty-weighted SDF / depth / colour losses and the pose-gradient chain.
// Replaces Mapper.sdf_losses == Tracker.sdf_losses (src/Mapper.py:141-175, src/Tracker.py:113-147), the
// loss assembly (Mapper.py:412-430, Tracker.py:208-228, 'original' mask modes), torch.median of the
// depth error (Tracker.py:213-215) and the autograd chain rays -> c2w -> (quaternion, translation)
// (src/common.py:102-105,196-208 + pytorch3d quaternion_to_matrix).
// Two-phase design: phase 1 accumulates sums and element counts (no boolean-index compaction, no
// host sync); phase 2 turns them into per-ray / per-sample upstream gradients.
#include "usl_loss.cuh"

namespace usl {

__global__ void __launch_bounds__(LOSS_WARPS * 32) loss_fwd_kernel(
    usl_loss_args_t a, const float *__restrict__ raw, const float *__restrict__ z, const float *__restrict__ gt_depth,
    const float *__restrict__ gt_color, const uint8_t *__restrict__ valid, const float *__restrict__ pixel_unc,
    const float *__restrict__ depth, const float *__restrict__ rgb, const float *__restrict__ median, int64_t R, int S,
    float *__restrict__ acc, uint8_t *__restrict__ mask_out) {
    __shared__ float s_acc[LOSS_WARPS][12];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t ray = (int64_t)blockIdx.x * LOSS_WARPS + warp;
    float v[12];
#pragma unroll
    for (int q = 0; q < 12; ++q) v[q] = 0.f;
    if (ray < R && (!valid || valid[ray])) {
        const float gt = gt_depth[ray], pu = pixel_unc[ray], dp = depth[ray];
        const bool m = ray_mask(a, gt, pu, dp, median);
        if (lane == 0 && mask_out) mask_out[ray] = m ? 1 : 0;
        const float tr = a.truncation, tr04 = a.truncation_center;
        if (m) {
            for (int s = lane; s < S; s += 32) {
                const float zz = z[ray * S + s], sd = raw[(ray * S + s) * 4 + 3];
                const int c = sample_class(zz, gt, tr, tr04);
                if (c == 0) { const float e = sd - 1.0f; v[A_FS] += e * e; v[N_FRONT] += 1.f; }
                else if (c < 3) {
                    const float e = (zz + sd * tr) - gt;
                    if (c == 1) { v[A_CENTER] += e * e; v[N_CENTER] += 1.f; }
                    else { v[A_TAIL] += e * e; v[N_TAIL] += 1.f; }
                }
            }
        }
        if (lane == 0) {
            v[N_RAYS] = 1.f;
            v[A_PUNC] = pu;
            if (m) { const float e = gt - dp; v[A_DEPTH] = e * e; v[N_MASK] = 1.f; }
            if (a.mode == 0 || m) {                                           // Mapper.py:427 (all rays) vs Tracker.py:225 (masked)
                float cs = 0.f;
#pragma unroll
                for (int k = 0; k < 3; ++k) { const float e = gt_color[ray * 3 + k] - rgb[ray * 3 + k]; cs += e * e; }
                v[A_COLOR] = cs; v[N_COLOR] = 3.f;
            }
        }
    } else if (ray < R && lane == 0 && mask_out) {
        mask_out[ray] = 0;
    }
#pragma unroll
    for (int q = 0; q < 12; ++q) v[q] = warp_sum(v[q]);
    if (lane == 0) {
#pragma unroll
        for (int q = 0; q < 12; ++q) s_acc[warp][q] = v[q];
    }
    __syncthreads();
    if (threadIdx.x < 12) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < LOSS_WARPS; ++w) s += s_acc[w][threadIdx.x];
        if (s != 0.f) atomicAdd(acc + threadIdx.x, s);
    }
}

__global__ void loss_finalize_kernel(usl_loss_args_t a, const float *__restrict__ acc, float *__restrict__ loss) {
    if (threadIdx.x == 0 && blockIdx.x == 0) loss[0] = loss_value(a, acc);
}

__global__ void __launch_bounds__(LOSS_WARPS * 32) loss_bwd_kernel(
    usl_loss_args_t a, const float *__restrict__ raw, const float *__restrict__ z, const float *__restrict__ gt_depth,
    const float *__restrict__ gt_color, const uint8_t *__restrict__ valid, const uint8_t *__restrict__ mask,
    const float *__restrict__ depth, const float *__restrict__ rgb, const float *__restrict__ acc,
    const float *__restrict__ g_loss, int64_t R, int S, float *__restrict__ g_depth, float *__restrict__ g_rgb,
    float *__restrict__ g_sdf) {
    const int lane = threadIdx.x & 31;
    const int64_t ray = (int64_t)blockIdx.x * LOSS_WARPS + (threadIdx.x >> 5);
    if (ray >= R) return;
    const float gl = g_loss ? g_loss[0] : 1.0f;
    const bool ok = !valid || valid[ray];
    const bool m = ok && mask[ray];
    const float gt = gt_depth[ray];
    const float tr = a.truncation, tr04 = a.truncation_center;
    const float k_fs = gl * a.w_sdf_fs * 2.0f / acc[N_FRONT];
    const float k_ce = gl * a.w_sdf_center * 2.0f * tr / acc[N_CENTER];
    const float k_ta = gl * a.w_sdf_tail * 2.0f * tr / acc[N_TAIL];
    for (int s = lane; s < S; s += 32) {
        float g = 0.f;
        if (m) {
            const float zz = z[ray * S + s], sd = raw[(ray * S + s) * 4 + 3];
            const int c = sample_class(zz, gt, tr, tr04);
            if (c == 0) g = k_fs * (sd - 1.0f);
            else if (c == 1) g = k_ce * ((zz + sd * tr) - gt);
            else if (c == 2) g = k_ta * ((zz + sd * tr) - gt);
        }
        g_sdf[ray * S + s] = g;
    }
    if (lane == 0) g_depth[ray] = m ? gl * a.w_depth * 2.0f * (depth[ray] - gt) / acc[N_MASK] : 0.f;
    if (lane < 3) {
        const bool on = ok && (a.mode == 0 || m);
        g_rgb[ray * 3 + lane] = on ? gl * a.w_color * 2.0f * (rgb[ray * 3 + lane] - gt_color[ray * 3 + lane]) / acc[N_COLOR] : 0.f;
    }
}

// ---- torch.median (lower middle) of |gt - depth| over valid rays: single-CTA radix select ----
#define MED_THREADS 1024
__global__ void __launch_bounds__(MED_THREADS) depth_error_median_kernel(const float *__restrict__ gt_depth,
                                                                         const float *__restrict__ depth,
                                                                         const uint8_t *__restrict__ valid, int64_t R,
                                                                         float *__restrict__ ws, float *__restrict__ median) {
    __shared__ unsigned int hist[256];
    __shared__ unsigned int s_count, s_prefix, s_k, s_nan;
    const int t = threadIdx.x;
    if (t == 0) { s_count = 0; s_nan = 0; }
    __syncthreads();
    // compact the errors of valid rays (order is irrelevant for a median)
    for (int64_t i = t; i < R; i += MED_THREADS) {
        if (!valid || valid[i]) {
            const float e = fabsf(gt_depth[i] - depth[i]);
            if (e != e) atomicAdd(&s_nan, 1u);
            ws[atomicAdd(&s_count, 1u)] = e;
        }
    }
    __syncthreads();
    const unsigned int n = s_count;
    if (n == 0 || s_nan) { if (t == 0) median[0] = NAN; return; }      // torch: median of empty / with NaN -> NaN
    if (t == 0) { s_prefix = 0; s_k = (n - 1) / 2; }                     // lower middle
    const unsigned int *bits = reinterpret_cast<const unsigned int *>(ws);
    for (int shift = 24; shift >= 0; shift -= 8) {
        if (t < 256) hist[t] = 0;
        __syncthreads();
        const unsigned int prefix = s_prefix;
        const unsigned int himask = (shift == 24) ? 0u : (0xFFFFFFFFu << (shift + 8));
        for (unsigned int i = t; i < n; i += MED_THREADS) {
            const unsigned int b = bits[i];                              // non-negative floats: bit pattern is order preserving
            if ((b & himask) == prefix) atomicAdd(&hist[(b >> shift) & 255u], 1u);
        }
        __syncthreads();
        if (t < 32) {                                    // warp 0: lane l owns bins 8l..8l+7, exclusive scan of the lane sums
            unsigned int own = 0;
#pragma unroll
            for (int q = 0; q < 8; ++q) own += hist[t * 8 + q];
            unsigned int inc = own;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned int v = __shfl_up_sync(0xffffffffu, inc, o);
                if (t >= o) inc += v;
            }
            const unsigned int k = s_k, before = inc - own;
            __syncwarp();                                 // every lane has read s_k before one of them replaces it
            if (k >= before && k < inc) {                 // exactly one lane: the k-th element falls into its bins
                unsigned int c = before;
                int bin = t * 8;
                for (; bin < t * 8 + 8; ++bin) { if (c + hist[bin] > k) break; c += hist[bin]; }
                s_k = k - c;
                s_prefix = prefix | ((unsigned int)bin << shift);
            }
        }
        __syncthreads();
    }
    if (t == 0) median[0] = __uint_as_float(s_prefix);
}

// Tracker.py:346-348: keep the pose at which the minimal loss was evaluated
__global__ void track_keep_best_kernel(const float *__restrict__ loss, const float *__restrict__ pose,
                                       float *__restrict__ best_loss, float *__restrict__ best_pose) {
    const bool better = loss[0] < best_loss[0];          // false for a NaN loss, as in the reference
    __syncwarp();
    if (better) {
        if (threadIdx.x < 7) best_pose[threadIdx.x] = pose[threadIdx.x];
        __syncwarp();
        if (threadIdx.x == 0) best_loss[0] = loss[0];
    }
}

// ---- pose gradient ------------------------------------------------------------------------------
// d_c2w[k][a*4+b] += d_rays_d[a]*dir_cam[b] (b<3) ; d_c2w[k][a*4+3] += d_rays_o[a]
__global__ void __launch_bounds__(256) pose_reduce_kernel(const float *__restrict__ d_o, const float *__restrict__ d_d,
                                                          const float *__restrict__ dirs, const int32_t *__restrict__ frame_id,
                                                          const uint8_t *__restrict__ valid, int64_t n, int K,
                                                          float *__restrict__ d_c2w) {
    const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    float v[12];
#pragma unroll
    for (int q = 0; q < 12; ++q) v[q] = 0.f;
    int fr = -1;
    if (m < n) {
        fr = frame_id ? frame_id[m] : 0;
        if (!valid || valid[m]) {
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                const float gd = d_d[m * 3 + a];
#pragma unroll
                for (int b = 0; b < 3; ++b) v[a * 4 + b] = gd * dirs[m * 3 + b];
                v[a * 4 + 3] = d_o[m * 3 + a];
            }
        }
    }
    const int fr0 = __shfl_sync(0xffffffffu, fr, 0);
    const bool uniform = __all_sync(0xffffffffu, fr == fr0 || fr < 0);
    if (uniform) {
#pragma unroll
        for (int q = 0; q < 12; ++q) v[q] = warp_sum(v[q]);
        if (lane == 0 && fr0 >= 0 && fr0 < K) {
#pragma unroll
            for (int q = 0; q < 12; ++q) if (v[q] != 0.f) atomicAdd(d_c2w + (int64_t)fr0 * 12 + q, v[q]);
        }
    } else if (fr >= 0 && fr < K) {
#pragma unroll
        for (int q = 0; q < 12; ++q) if (v[q] != 0.f) atomicAdd(d_c2w + (int64_t)fr * 12 + q, v[q]);
    }
}

// pytorch3d quaternion_to_matrix (real first, not normalised) + translation -> 4x4
__global__ void pose_to_matrix_kernel(const float *__restrict__ pose, int K, float *__restrict__ c2w) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= K) return;
    float m[12];
    pose_to_c2w(pose + k * 7, m);
    float *o = c2w + k * 16;
#pragma unroll
    for (int q = 0; q < 12; ++q) o[q] = m[q];
    o[12] = 0.f; o[13] = 0.f; o[14] = 0.f; o[15] = 1.f;
}

// R_ab = delta_ab + s2 * M_ab(q), s2 = 2/|q|^2  =>  dL/dq_m = s2 * sum G_ab dM_ab/dq_m - s2^2 q_m sum G_ab M_ab
__global__ void pose_matrix_bwd_kernel(const float *__restrict__ pose, const float *__restrict__ d_c2w, int K,
                                       float *__restrict__ d_pose) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= K) return;
    const float r = pose[k * 7], i = pose[k * 7 + 1], j = pose[k * 7 + 2], q = pose[k * 7 + 3];
    const float *G = d_c2w + k * 12;
    const float G00 = G[0], G01 = G[1], G02 = G[2], G10 = G[4], G11 = G[5], G12 = G[6], G20 = G[8], G21 = G[9], G22 = G[10];
    const float n = r * r + i * i + j * j + q * q;
    const float s2 = 2.0f / n;
    const float M00 = -(j * j + q * q), M01 = i * j - q * r, M02 = i * q + j * r;
    const float M10 = i * j + q * r, M11 = -(i * i + q * q), M12 = j * q - i * r;
    const float M20 = i * q - j * r, M21 = j * q + i * r, M22 = -(i * i + j * j);
    const float GM = G00 * M00 + G01 * M01 + G02 * M02 + G10 * M10 + G11 * M11 + G12 * M12 + G20 * M20 + G21 * M21 + G22 * M22;
    const float dr = -q * G01 + j * G02 + q * G10 - i * G12 - j * G20 + i * G21;
    const float di = j * G01 + q * G02 + j * G10 - 2 * i * G11 - r * G12 + q * G20 + r * G21 - 2 * i * G22;
    const float dj = -2 * j * G00 + i * G01 + r * G02 + i * G10 + q * G12 - r * G20 + q * G21 - 2 * j * G22;
    const float dq = -2 * q * G00 - r * G01 + i * G02 + r * G10 - 2 * q * G11 + j * G12 + i * G20 + j * G21;
    const float c = s2 * s2 * GM;
    d_pose[k * 7 + 0] = s2 * dr - c * r;
    d_pose[k * 7 + 1] = s2 * di - c * i;
    d_pose[k * 7 + 2] = s2 * dj - c * j;
    d_pose[k * 7 + 3] = s2 * dq - c * q;
    d_pose[k * 7 + 4] = G[3]; d_pose[k * 7 + 5] = G[7]; d_pose[k * 7 + 6] = G[11];
}

}  // namespace usl

using namespace usl;

extern "C" {

int usl_loss_fwd(const usl_loss_args_t *a, const float *raw, const float *z, const float *gt_depth,
                 const float *gt_color, const uint8_t *valid, const float *pixel_unc, const float *depth,
                 const float *rgb, const float *median, int64_t R, int S, float *acc, uint8_t *mask_out,
                 usl_stream_t stream) {
    if (R <= 0) return 0;
    if (!a || (a->mode == 1 && !median)) { set_error("usl_loss_fwd: tracking mode needs the depth-error median"); return 1; }
    loss_fwd_kernel<<<(unsigned)((R + LOSS_WARPS - 1) / LOSS_WARPS), LOSS_WARPS * 32, 0, (cudaStream_t)stream>>>(
        *a, raw, z, gt_depth, gt_color, valid, pixel_unc, depth, rgb, median, R, S, acc, mask_out);
    return check_launch("usl_loss_fwd");
}

int usl_loss_finalize(const usl_loss_args_t *a, const float *acc, float *loss, usl_stream_t stream) {
    loss_finalize_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(*a, acc, loss);
    return check_launch("usl_loss_finalize");
}

int usl_loss_bwd(const usl_loss_args_t *a, const float *raw, const float *z, const float *gt_depth,
                 const float *gt_color, const uint8_t *valid, const uint8_t *mask, const float *depth,
                 const float *rgb, const float *acc, const float *g_loss, int64_t R, int S, float *g_depth,
                 float *g_rgb, float *g_sdf, usl_stream_t stream) {
    if (R <= 0) return 0;
    loss_bwd_kernel<<<(unsigned)((R + LOSS_WARPS - 1) / LOSS_WARPS), LOSS_WARPS * 32, 0, (cudaStream_t)stream>>>(
        *a, raw, z, gt_depth, gt_color, valid, mask, depth, rgb, acc, g_loss, R, S, g_depth, g_rgb, g_sdf);
    return check_launch("usl_loss_bwd");
}

int usl_depth_error_median(const float *gt_depth, const float *depth, const uint8_t *valid, int64_t R,
                           float *workspace, float *median, usl_stream_t stream) {
    if (R <= 0) { set_error("usl_depth_error_median: empty input"); return 1; }
    depth_error_median_kernel<<<1, MED_THREADS, 0, (cudaStream_t)stream>>>(gt_depth, depth, valid, R, workspace, median);
    return check_launch("usl_depth_error_median");
}

int usl_track_keep_best(const float *loss, const float *cam_pose, float *best_loss, float *best_pose, usl_stream_t stream) {
    if (!loss || !cam_pose || !best_loss || !best_pose) { set_error("usl_track_keep_best: null argument"); return 1; }
    track_keep_best_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(loss, cam_pose, best_loss, best_pose);
    return check_launch("usl_track_keep_best");
}

int usl_pose_reduce(const float *d_rays_o, const float *d_rays_d, const float *dirs_cam, const int32_t *frame_id,
                    const uint8_t *valid, int64_t n, int K, float *d_c2w, usl_stream_t stream) {
    if (n <= 0) return 0;
    pose_reduce_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(d_rays_o, d_rays_d, dirs_cam, frame_id, valid, n, K, d_c2w);
    return check_launch("usl_pose_reduce");
}

int usl_pose_to_matrix(const float *pose, int K, float *c2w, usl_stream_t stream) {
    if (K <= 0) return 0;
    pose_to_matrix_kernel<<<(K + 127) / 128, 128, 0, (cudaStream_t)stream>>>(pose, K, c2w);
    return check_launch("usl_pose_to_matrix");
}

int usl_pose_matrix_bwd(const float *pose, const float *d_c2w, int K, float *d_pose, usl_stream_t stream) {
    if (K <= 0) return 0;
    pose_matrix_bwd_kernel<<<(K + 127) / 128, 128, 0, (cudaStream_t)stream>>>(pose, d_c2w, K, d_pose);
    return check_launch("usl_pose_matrix_bwd");
}

}  // extern "C"
