// SDF -> alpha -> transmittance -> weights compositing of depth / colour / uncertainty, one warp per
// ray with shuffle scans, forward and backward.  Replaces src/utils/Renderer.py:140-158 (sdf2alpha,
// cumprod transmittance, weighted sums) and the ~25 autograd nodes behind it.
#include "usl_loss.cuh"

namespace usl {

#define CMP_MAX_CHUNKS 4   // S <= 128
#define CMP_WARPS 8

struct Sample {
    float sdf, z, alpha, T, w, sg;   // sg = sigmoid(-sdf*beta)
    float c[3];
};

// inclusive warp product scan
__device__ __forceinline__ float warp_scan_mul(float v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const float t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v *= t;
    }
    return v;
}
// inclusive warp suffix sum (reverse scan)
__device__ __forceinline__ float warp_rscan_add(float v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const float t = __shfl_down_sync(0xffffffffu, v, o);
        if (lane + o < 32) v += t;
    }
    return v;
}

// Loads the ray's samples and runs the forward scan. Returns sums via references.
__device__ __forceinline__ void ray_forward(const float *__restrict__ raw, const float *__restrict__ z, float beta,
                                            int64_t ray, int S, int lane, Sample smp[CMP_MAX_CHUNKS], float &term,
                                            float &depth, float rgb[3]) {
    float carry = 1.0f, sw = 0.f, swz = 0.f, sc0 = 0.f, sc1 = 0.f, sc2 = 0.f;
#pragma unroll
    for (int ch = 0; ch < CMP_MAX_CHUNKS; ++ch) {
        const int s = ch * 32 + lane;
        Sample &q = smp[ch];
        if (ch * 32 >= S) { q.w = 0.f; q.alpha = 0.f; q.T = 0.f; q.z = 0.f; q.sdf = 0.f; q.sg = 0.f; q.c[0] = q.c[1] = q.c[2] = 0.f; continue; }
        const bool in = s < S;
        float4 rv = make_float4(0.f, 0.f, 0.f, 0.f);
        float zz = 0.f;
        if (in) {
            rv = __ldg(reinterpret_cast<const float4 *>(raw) + ray * S + s);
            zz = __ldg(z + ray * S + s);
        }
        q.sdf = rv.w; q.z = zz; q.c[0] = rv.x; q.c[1] = rv.y; q.c[2] = rv.z;
        q.sg = 1.0f / (1.0f + expf(rv.w * beta));
        q.alpha = in ? 1.0f - expf(-beta * q.sg) : 0.f;
        const float p = in ? (1.0f - q.alpha) + 1e-10f : 1.0f;
        const float incl = warp_scan_mul(p, lane);
        float excl = __shfl_up_sync(0xffffffffu, incl, 1);
        if (lane == 0) excl = 1.0f;
        q.T = carry * excl;
        q.w = q.alpha * q.T;
        carry *= __shfl_sync(0xffffffffu, incl, 31);
        sw += q.w; swz += q.w * zz;
        sc0 += q.w * rv.x; sc1 += q.w * rv.y; sc2 += q.w * rv.z;
    }
    term = warp_sum(sw);
    depth = warp_sum(swz);
    rgb[0] = warp_sum(sc0); rgb[1] = warp_sum(sc1); rgb[2] = warp_sum(sc2);
}

__global__ void __launch_bounds__(CMP_WARPS * 32) composite_fwd_kernel(
    const float *__restrict__ raw, const float *__restrict__ z, const float *__restrict__ beta_p,
    const uint8_t *__restrict__ valid, int64_t R, int S, float *__restrict__ term_o, float *__restrict__ punc_o,
    float *__restrict__ depth_o, float *__restrict__ rgb_o, float *__restrict__ dunc_o, float *__restrict__ weights_o) {
    const int lane = threadIdx.x & 31;
    const int64_t ray = (int64_t)blockIdx.x * CMP_WARPS + (threadIdx.x >> 5);
    if (ray >= R) return;
    if (valid && !valid[ray]) {
        if (lane == 0) {
            term_o[ray] = 0.f; punc_o[ray] = 0.f; depth_o[ray] = 0.f; dunc_o[ray] = 0.f;
            rgb_o[ray * 3] = 0.f; rgb_o[ray * 3 + 1] = 0.f; rgb_o[ray * 3 + 2] = 0.f;
        }
        if (weights_o) for (int s = lane; s < S; s += 32) weights_o[ray * S + s] = 0.f;
        return;
    }
    const float beta = beta_p[0];
    Sample smp[CMP_MAX_CHUNKS];
    float term, depth, rgb[3];
    ray_forward(raw, z, beta, ray, S, lane, smp, term, depth, rgb);
    float v = 0.f;
#pragma unroll
    for (int ch = 0; ch < CMP_MAX_CHUNKS; ++ch) {
        const float dz = depth - smp[ch].z;
        v += smp[ch].w * dz * dz;
        if (weights_o && ch * 32 + lane < S) weights_o[ray * S + ch * 32 + lane] = smp[ch].w;
    }
    v = warp_sum(v);
    if (lane == 0) {
        term_o[ray] = term;
        const float om = 1.0f - term;
        punc_o[ray] = om * om;                               // Renderer.py:148
        depth_o[ray] = depth;
        dunc_o[ray] = sqrtf(v);                              // Renderer.py:150
        rgb_o[ray * 3] = rgb[0]; rgb_o[ray * 3 + 1] = rgb[1]; rgb_o[ray * 3 + 2] = rgb[2];
    }
}

// composite_fwd + loss_fwd for the modes whose ray mask needs no global quantity (mapping 'original', 'no_mask'): the warp that
// composited a ray still holds its samples, so the loss sums cost no second pass over raw / z and no second launch.
struct CompLossFwdArgs {
    const float *raw, *z, *beta, *gt_depth, *gt_color;
    const uint8_t *valid;
    int64_t R;
    int S;
    float *term, *punc, *depth, *rgb, *dunc, *acc;
    uint8_t *mask_out;
    usl_loss_args_t la;
};

__global__ void __launch_bounds__(CMP_WARPS * 32) composite_loss_fwd_kernel(const __grid_constant__ CompLossFwdArgs A) {
    __shared__ float s_acc[CMP_WARPS][12];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t ray = (int64_t)blockIdx.x * CMP_WARPS + warp;
    const int S = A.S;
    float v[12];
#pragma unroll
    for (int q = 0; q < 12; ++q) v[q] = 0.f;
    if (ray < A.R && A.valid && !A.valid[ray]) {
        if (lane == 0) {
            A.term[ray] = 0.f; A.punc[ray] = 0.f; A.depth[ray] = 0.f; A.dunc[ray] = 0.f;
            A.rgb[ray * 3] = 0.f; A.rgb[ray * 3 + 1] = 0.f; A.rgb[ray * 3 + 2] = 0.f;
            if (A.mask_out) A.mask_out[ray] = 0;
        }
    } else if (ray < A.R) {
        const float beta = A.beta[0];
        Sample smp[CMP_MAX_CHUNKS];
        float term, depth, rgb[3];
        ray_forward(A.raw, A.z, beta, ray, S, lane, smp, term, depth, rgb);
        float dv = 0.f;
#pragma unroll
        for (int ch = 0; ch < CMP_MAX_CHUNKS; ++ch) { const float dz = depth - smp[ch].z; dv += smp[ch].w * dz * dz; }
        dv = warp_sum(dv);
        const float om = 1.0f - term, pu = om * om;                // Renderer.py:148
        const float gt = A.gt_depth[ray];
        const bool m = ray_mask(A.la, gt, pu, depth, nullptr);
        if (m) {                                                   // same per-lane accumulation order as loss_fwd_kernel
            const float tr = A.la.truncation, tr04 = A.la.truncation_center;
#pragma unroll
            for (int ch = 0; ch < CMP_MAX_CHUNKS; ++ch) {
                if (ch * 32 + lane >= S) continue;
                const float zz = smp[ch].z, sd = smp[ch].sdf;
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
            A.term[ray] = term; A.punc[ray] = pu; A.depth[ray] = depth;
            A.dunc[ray] = sqrtf(dv);                               // Renderer.py:150
            A.rgb[ray * 3] = rgb[0]; A.rgb[ray * 3 + 1] = rgb[1]; A.rgb[ray * 3 + 2] = rgb[2];
            if (A.mask_out) A.mask_out[ray] = m ? 1 : 0;
            v[N_RAYS] = 1.f;
            v[A_PUNC] = pu;
            if (m) { const float e = gt - depth; v[A_DEPTH] = e * e; v[N_MASK] = 1.f; }
            if (A.la.mode == 0 || m) {                             // Mapper.py:427 (all rays) / no_mask
                float cs = 0.f;
#pragma unroll
                for (int k = 0; k < 3; ++k) { const float e = A.gt_color[ray * 3 + k] - rgb[k]; cs += e * e; }
                v[A_COLOR] = cs; v[N_COLOR] = 3.f;
            }
        }
    }
#pragma unroll
    for (int q = 0; q < 12; ++q) v[q] = warp_sum(v[q]);
    if (lane == 0) {
#pragma unroll
        for (int q = 0; q < 12; ++q) s_acc[warp][q] = v[q];
    }
    __syncthreads();
    if (threadIdx.x < 12) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < CMP_WARPS; ++w) t += s_acc[w][threadIdx.x];
        if (t != 0.f) atomicAdd(A.acc + threadIdx.x, t);
    }
}

struct CompBwdArgs {
    const float *raw, *z, *beta;
    const uint8_t *valid;
    int64_t R;
    int S;
    const float *g_term, *g_punc, *g_depth, *g_rgb, *g_dunc, *g_sdf, *jac;
    usl_bound_t bound;
    float *d_raw, *d_beta, *d_rays_o, *d_rays_d;
    // fused loss gradient (usl_composite_loss_bwd): upstream gradients derived in place from the loss definition
    int fused_loss;
    usl_loss_args_t la;
    const float *gt_depth, *gt_color, *depth, *rgb, *acc, *g_loss;
    const uint8_t *mask;
    float *loss;
};

__global__ void __launch_bounds__(CMP_WARPS * 32) composite_bwd_kernel(const __grid_constant__ CompBwdArgs A) {
    __shared__ float s_dbeta[CMP_WARPS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t ray = (int64_t)blockIdx.x * CMP_WARPS + warp;
    const int S = A.S;
    float dbeta = 0.f;
    const bool live = ray < A.R;
    if (live && A.valid && !A.valid[ray]) {
        for (int s = lane; s < S; s += 32) reinterpret_cast<float4 *>(A.d_raw)[ray * S + s] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (A.jac && lane < 3) { A.d_rays_o[ray * 3 + lane] = 0.f; A.d_rays_d[ray * 3 + lane] = 0.f; }
    } else if (live) {
        if (A.jac) {
            // the Jacobian rows are only needed after two dependent scans: start them towards L1 now, so the kernel pays one
            // memory round trip (raw / z, below) instead of two in sequence
            const int64_t npts = A.R * (int64_t)S;
            for (int s = lane; s < S; s += 32) {
                const float *jp = A.jac + ray * S + s;
#pragma unroll
                for (int c = 0; c < 12; ++c) asm volatile("prefetch.global.L1 [%0];" ::"l"(jp + (int64_t)c * npts));
            }
        }
        const float beta = A.beta[0];
        Sample smp[CMP_MAX_CHUNKS];
        float term, depth, rgb[3];
        ray_forward(A.raw, A.z, beta, ray, S, lane, smp, term, depth, rgb);
        // upstream gradients of the per-ray outputs
        float g_depth = A.g_depth ? A.g_depth[ray] : 0.f;
        float g_term = A.g_term ? A.g_term[ray] : 0.f;
        if (A.g_punc) g_term += A.g_punc[ray] * (-2.0f * (1.0f - term));
        float g_c[3] = {0.f, 0.f, 0.f};
        if (A.g_rgb) { g_c[0] = A.g_rgb[ray * 3]; g_c[1] = A.g_rgb[ray * 3 + 1]; g_c[2] = A.g_rgb[ray * 3 + 2]; }
        // fused loss: same formulas as loss_bwd_kernel (Mapper.py:141-175,422-430 / Tracker.py:220-228)
        bool lmask = false;
        float gt = 0.f, k_fs = 0.f, k_ce = 0.f, k_ta = 0.f;
        if (A.fused_loss) {
            const float gl = A.g_loss ? A.g_loss[0] : 1.0f;
            lmask = A.mask[ray] != 0;
            gt = A.gt_depth[ray];
            const float tr = A.la.truncation;
            k_fs = gl * A.la.w_sdf_fs * 2.0f / A.acc[N_FRONT];
            k_ce = gl * A.la.w_sdf_center * 2.0f * tr / A.acc[N_CENTER];
            k_ta = gl * A.la.w_sdf_tail * 2.0f * tr / A.acc[N_TAIL];
            g_depth = lmask ? gl * A.la.w_depth * 2.0f * (A.depth[ray] - gt) / A.acc[N_MASK] : 0.f;
            const bool con = (A.la.mode == 0) || lmask;
#pragma unroll
            for (int k = 0; k < 3; ++k)
                g_c[k] = con ? gl * A.la.w_color * 2.0f * (A.rgb[ray * 3 + k] - A.gt_color[ray * 3 + k]) / A.acc[N_COLOR] : 0.f;
        }
        float dV = 0.f;
        const float g_dunc = A.g_dunc ? A.g_dunc[ray] : 0.f;
        if (g_dunc != 0.f) {
            float v = 0.f;
#pragma unroll
            for (int ch = 0; ch < CMP_MAX_CHUNKS; ++ch) { const float dz = depth - smp[ch].z; v += smp[ch].w * dz * dz; }
            v = warp_sum(v);
            dV = g_dunc / (2.0f * sqrtf(v));
            g_depth += dV * 2.0f * depth * (term - 1.0f);   // d/d depth of sum w (depth - z)^2 = 2 (depth*term - depth)
        }
        // g_w per sample, then suffix sums of g_w*w for the transmittance chain (processed back to front)
        float gw[CMP_MAX_CHUNKS];
#pragma unroll
        for (int ch = 0; ch < CMP_MAX_CHUNKS; ++ch) {
            const Sample &q = smp[ch];
            const float dz = depth - q.z;
            gw[ch] = g_depth * q.z + g_c[0] * q.c[0] + g_c[1] * q.c[1] + g_c[2] * q.c[2] + g_term + dV * dz * dz;
        }
        float carry = 0.f;
        float go[3] = {0.f, 0.f, 0.f}, gd[3] = {0.f, 0.f, 0.f};
#pragma unroll
        for (int ch = CMP_MAX_CHUNKS - 1; ch >= 0; --ch) {
            if (ch * 32 >= S) continue;
            const Sample &q = smp[ch];
            const int s = ch * 32 + lane;
            const bool in = s < S;
            const float prod = in ? gw[ch] * q.w : 0.f;
            const float incl = warp_rscan_add(prod, lane);
            const float suffix = carry + (incl - prod);                 // sum over k > s
            carry += __shfl_sync(0xffffffffu, incl, 0);
            const float one_m = (1.0f - q.alpha) + 1e-10f;
            const float g_alpha = gw[ch] * q.T - suffix / one_m;
            const float e = 1.0f - q.alpha;                             // exp(-beta*sg)
            const float dsg = q.sg * (1.0f - q.sg);
            // alpha = 1 - exp(-beta*sg), sg = sigmoid(-sdf*beta)
            float g_s = g_alpha * (beta * e) * (-beta * dsg);
            if (in) dbeta += g_alpha * (q.sg * e + beta * e * (-q.sdf * dsg));
            if (A.g_sdf && in) g_s += A.g_sdf[ray * S + s];
            if (A.fused_loss && lmask && in) {
                const int c = sample_class(q.z, gt, A.la.truncation, A.la.truncation_center);
                if (c == 0) g_s += k_fs * (q.sdf - 1.0f);
                else if (c == 1) g_s += k_ce * ((q.z + q.sdf * A.la.truncation) - gt);
                else if (c == 2) g_s += k_ta * ((q.z + q.sdf * A.la.truncation) - gt);
            }
            if (in) {
                const float4 dr = make_float4(q.w * g_c[0], q.w * g_c[1], q.w * g_c[2], g_s);
                reinterpret_cast<float4 *>(A.d_raw)[ray * S + s] = dr;
                if (A.jac) {
                    // jac is component-major [12][R*S] (rows r,g,b,sdf x 3 dims): every load is coalesced across the samples
                    const int64_t npts = A.R * (int64_t)S, pi = ray * S + s;
                    float gx0 = 0.f, gx1 = 0.f, gx2 = 0.f;
                    const float drv[4] = {dr.x, dr.y, dr.z, dr.w};
#pragma unroll
                    for (int o = 0; o < 4; ++o) {
                        gx0 = fmaf(drv[o], __ldg(A.jac + (int64_t)(o * 3 + 0) * npts + pi), gx0);
                        gx1 = fmaf(drv[o], __ldg(A.jac + (int64_t)(o * 3 + 1) * npts + pi), gx1);
                        gx2 = fmaf(drv[o], __ldg(A.jac + (int64_t)(o * 3 + 2) * npts + pi), gx2);
                    }
                    const float gp0 = gx0 / (A.bound.hi[0] - A.bound.lo[0]);
                    const float gp1 = gx1 / (A.bound.hi[1] - A.bound.lo[1]);
                    const float gp2 = gx2 / (A.bound.hi[2] - A.bound.lo[2]);
                    go[0] += gp0; go[1] += gp1; go[2] += gp2;
                    gd[0] += gp0 * q.z; gd[1] += gp1 * q.z; gd[2] += gp2 * q.z;
                }
            }
        }
        if (A.jac) {
#pragma unroll
            for (int d = 0; d < 3; ++d) { go[d] = warp_sum(go[d]); gd[d] = warp_sum(gd[d]); }
            if (lane == 0) {
#pragma unroll
                for (int d = 0; d < 3; ++d) { A.d_rays_o[ray * 3 + d] = go[d]; A.d_rays_d[ray * 3 + d] = gd[d]; }
            }
        }
        dbeta = warp_sum(dbeta);
    }
    if (A.fused_loss && A.loss && blockIdx.x == 0 && threadIdx.x == 0) A.loss[0] = loss_value(A.la, A.acc);   // == usl_loss_finalize
    if (A.d_beta) {
        if (lane == 0) s_dbeta[warp] = dbeta;
        __syncthreads();
        if (threadIdx.x == 0) {
            float s = 0.f;
#pragma unroll
            for (int w = 0; w < CMP_WARPS; ++w) s += s_dbeta[w];
            if (s != 0.f) atomicAdd(A.d_beta, s);
        }
    }
}

}  // namespace usl

using namespace usl;

extern "C" {

int usl_composite_fwd(const float *raw, const float *z, const float *beta, const uint8_t *valid, int64_t R, int S,
                      float *term, float *pixel_unc, float *depth, float *rgb, float *depth_unc, float *weights,
                      usl_stream_t stream) {
    if (R <= 0) return 0;
    if (S < 1 || S > CMP_MAX_CHUNKS * 32) { set_error("usl_composite_fwd: S must be in 1..128"); return 1; }
    composite_fwd_kernel<<<(unsigned)((R + CMP_WARPS - 1) / CMP_WARPS), CMP_WARPS * 32, 0, (cudaStream_t)stream>>>(
        raw, z, beta, valid, R, S, term, pixel_unc, depth, rgb, depth_unc, weights);
    return check_launch("usl_composite_fwd");
}

int usl_composite_loss_fwd(const usl_loss_args_t *a, const float *raw, const float *z, const float *beta, const uint8_t *valid,
                           int64_t R, int S, const float *gt_depth, const float *gt_color, float *term, float *pixel_unc,
                           float *depth, float *rgb, float *depth_unc, float *acc, uint8_t *mask_out, usl_stream_t stream) {
    if (R <= 0) return 0;
    if (S < 1 || S > CMP_MAX_CHUNKS * 32) { set_error("usl_composite_loss_fwd: S must be in 1..128"); return 1; }
    if (!a || !raw || !z || !beta || !gt_depth || !gt_color || !term || !pixel_unc || !depth || !rgb || !depth_unc || !acc) {
        set_error("usl_composite_loss_fwd: null argument"); return 1;
    }
    if (a->mode == 1) { set_error("usl_composite_loss_fwd: the tracking mask needs the median of all rays: use usl_composite_fwd + usl_depth_error_median + usl_loss_fwd"); return 1; }
    CompLossFwdArgs A;
    A.raw = raw; A.z = z; A.beta = beta; A.gt_depth = gt_depth; A.gt_color = gt_color; A.valid = valid; A.R = R; A.S = S;
    A.term = term; A.punc = pixel_unc; A.depth = depth; A.rgb = rgb; A.dunc = depth_unc; A.acc = acc; A.mask_out = mask_out; A.la = *a;
    composite_loss_fwd_kernel<<<(unsigned)((R + CMP_WARPS - 1) / CMP_WARPS), CMP_WARPS * 32, 0, (cudaStream_t)stream>>>(A);
    return check_launch("usl_composite_loss_fwd");
}

int usl_composite_bwd(const float *raw, const float *z, const float *beta, const uint8_t *valid, int64_t R, int S,
                      const float *g_term, const float *g_punc, const float *g_depth, const float *g_rgb,
                      const float *g_dunc, const float *g_sdf, const float *jac, const usl_bound_t *bound,
                      float *d_raw, float *d_beta, float *d_rays_o, float *d_rays_d, usl_stream_t stream) {
    if (R <= 0) return 0;
    if (S < 1 || S > CMP_MAX_CHUNKS * 32) { set_error("usl_composite_bwd: S must be in 1..128"); return 1; }
    if (jac && (!bound || !d_rays_o || !d_rays_d)) { set_error("usl_composite_bwd: jac needs bound, d_rays_o, d_rays_d"); return 1; }
    CompBwdArgs A;
    A.raw = raw; A.z = z; A.beta = beta; A.valid = valid; A.R = R; A.S = S;
    A.g_term = g_term; A.g_punc = g_punc; A.g_depth = g_depth; A.g_rgb = g_rgb; A.g_dunc = g_dunc; A.g_sdf = g_sdf; A.jac = jac;
    if (bound) A.bound = *bound; else { for (int d = 0; d < 3; ++d) { A.bound.lo[d] = 0.f; A.bound.hi[d] = 1.f; } }
    A.d_raw = d_raw; A.d_beta = d_beta; A.d_rays_o = d_rays_o; A.d_rays_d = d_rays_d;
    A.fused_loss = 0; A.loss = nullptr; A.mask = nullptr;
    composite_bwd_kernel<<<(unsigned)((R + CMP_WARPS - 1) / CMP_WARPS), CMP_WARPS * 32, 0, (cudaStream_t)stream>>>(A);
    return check_launch("usl_composite_bwd");
}

int usl_composite_loss_bwd(const usl_loss_args_t *a, const float *raw, const float *z, const float *beta,
                           const uint8_t *valid, const uint8_t *mask, int64_t R, int S, const float *gt_depth,
                           const float *gt_color, const float *depth, const float *rgb, const float *acc,
                           const float *g_loss, const float *jac, const usl_bound_t *bound, float *d_raw, float *d_beta,
                           float *d_rays_o, float *d_rays_d, float *loss, usl_stream_t stream) {
    if (R <= 0) return 0;
    if (S < 1 || S > CMP_MAX_CHUNKS * 32) { set_error("usl_composite_loss_bwd: S must be in 1..128"); return 1; }
    if (!a || !mask || !acc || !gt_depth || !gt_color || !depth || !rgb) { set_error("usl_composite_loss_bwd: null argument"); return 1; }
    if (jac && (!bound || !d_rays_o || !d_rays_d)) { set_error("usl_composite_loss_bwd: jac needs bound, d_rays_o, d_rays_d"); return 1; }
    CompBwdArgs A;
    A.raw = raw; A.z = z; A.beta = beta; A.valid = valid; A.R = R; A.S = S;
    A.g_term = nullptr; A.g_punc = nullptr; A.g_depth = nullptr; A.g_rgb = nullptr; A.g_dunc = nullptr; A.g_sdf = nullptr; A.jac = jac;
    if (bound) A.bound = *bound; else { for (int d = 0; d < 3; ++d) { A.bound.lo[d] = 0.f; A.bound.hi[d] = 1.f; } }
    A.d_raw = d_raw; A.d_beta = d_beta; A.d_rays_o = d_rays_o; A.d_rays_d = d_rays_d;
    A.fused_loss = 1; A.la = *a; A.gt_depth = gt_depth; A.gt_color = gt_color; A.depth = depth; A.rgb = rgb; A.acc = acc;
    A.g_loss = g_loss; A.mask = mask; A.loss = loss;
    composite_bwd_kernel<<<(unsigned)((R + CMP_WARPS - 1) / CMP_WARPS), CMP_WARPS * 32, 0, (cudaStream_t)stream>>>(A);
    return check_launch("usl_composite_loss_bwd");
}

}  // extern "C"
