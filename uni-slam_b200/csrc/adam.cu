// f1: fused multi-tensor Adam with optional gradient zeroing, same update rule as the host code's
// torch.optim.Adam(optimizer_config) (src/Mapper.py:358-364,445; src/Tracker.py:324-329,242): no weight decay,
// no amsgrad, per-group lr / betas / eps.  One launch over every parameter (tables 12.9 M floats + decoders +
// beta + poses) instead of torch's multi-pass foreach implementation (~28 B/param of traffic -> 32 B/param once).
#include <cmath>

#include "usl_device.cuh"

namespace usl {

struct AdamArgs {
    usl_adam_group_t g[USL_ADAM_MAX_GROUPS];
    int32_t first_block[USL_ADAM_MAX_GROUPS + 1];
    float step_size[USL_ADAM_MAX_GROUPS];   // lr / (1 - beta1^t)            (host double math, as torch does)
    float bc2_sqrt[USL_ADAM_MAX_GROUPS];    // sqrt(1 - beta2^t)
    int32_t n_groups;
    int32_t zero_grad;
    const int64_t *step_dev;                // optional device-side step counter (CUDA-graph replay)
};

#define ADAM_THREADS 256
#define ADAM_VEC_PER_THREAD 4               // float4 x 4 = 16 floats per thread, 4096 per CTA

__device__ __forceinline__ void adam_update(float &p, float &m, float &v, float g, float b1, float b2, float eps,
                                            float step_size, float bc2s) {
    m = m + (1.0f - b1) * (g - m);                       // exp_avg.lerp_(grad, 1 - beta1), weight < 0.5 branch
    v = v * b2 + (1.0f - b2) * g * g;                    // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value=1-beta2)
    const float denom = sqrtf(v) / bc2s + eps;           // (exp_avg_sq.sqrt() / bias_correction2_sqrt).add_(eps)
    p = p - step_size * (m / denom);                     // param.addcdiv_(exp_avg, denom, value=-step_size)
}

__global__ void __launch_bounds__(ADAM_THREADS) adam_kernel(const __grid_constant__ AdamArgs A) {
    __shared__ float s_step_size, s_bc2s;
    int gi = 0;
    while (gi + 1 < A.n_groups && (int)blockIdx.x >= A.first_block[gi + 1]) ++gi;
    const usl_adam_group_t &G = A.g[gi];
    float step_size = A.step_size[gi], bc2s = A.bc2_sqrt[gi];
    if (A.step_dev) {
        if (threadIdx.x == 0) {
            const double t = (double)A.step_dev[0];
            s_step_size = (float)((double)G.lr / (1.0 - pow((double)G.beta1, t)));
            s_bc2s = (float)sqrt(1.0 - pow((double)G.beta2, t));
        }
        __syncthreads();
        step_size = s_step_size; bc2s = s_bc2s;
    }
    const int64_t base = (int64_t)(blockIdx.x - A.first_block[gi]) * (ADAM_THREADS * ADAM_VEC_PER_THREAD * 4);
    const int64_t n = G.n;
    const bool vec_ok = ((reinterpret_cast<uintptr_t>(G.param) | reinterpret_cast<uintptr_t>(G.grad) |
                          reinterpret_cast<uintptr_t>(G.exp_avg) | reinterpret_cast<uintptr_t>(G.exp_avg_sq)) & 15u) == 0;
#pragma unroll
    for (int q = 0; q < ADAM_VEC_PER_THREAD; ++q) {
        const int64_t i = base + ((int64_t)q * ADAM_THREADS + threadIdx.x) * 4;
        if (i >= n) break;
        if (vec_ok && i + 4 <= n) {
            float4 p = *reinterpret_cast<float4 *>(G.param + i);
            const float4 g = *reinterpret_cast<const float4 *>(G.grad + i);
            float4 m = *reinterpret_cast<float4 *>(G.exp_avg + i);
            float4 v = *reinterpret_cast<float4 *>(G.exp_avg_sq + i);
            adam_update(p.x, m.x, v.x, g.x, G.beta1, G.beta2, G.eps, step_size, bc2s);
            adam_update(p.y, m.y, v.y, g.y, G.beta1, G.beta2, G.eps, step_size, bc2s);
            adam_update(p.z, m.z, v.z, g.z, G.beta1, G.beta2, G.eps, step_size, bc2s);
            adam_update(p.w, m.w, v.w, g.w, G.beta1, G.beta2, G.eps, step_size, bc2s);
            *reinterpret_cast<float4 *>(G.param + i) = p;
            *reinterpret_cast<float4 *>(G.exp_avg + i) = m;
            *reinterpret_cast<float4 *>(G.exp_avg_sq + i) = v;
            if (A.zero_grad) *reinterpret_cast<float4 *>(G.grad + i) = make_float4(0.f, 0.f, 0.f, 0.f);
        } else {
            for (int64_t j = i; j < n && j < i + 4; ++j) {
                float p = G.param[j], m = G.exp_avg[j], v = G.exp_avg_sq[j];
                adam_update(p, m, v, G.grad[j], G.beta1, G.beta2, G.eps, step_size, bc2s);
                G.param[j] = p; G.exp_avg[j] = m; G.exp_avg_sq[j] = v;
                if (A.zero_grad) G.grad[j] = 0.f;
            }
        }
    }
}

}  // namespace usl

using namespace usl;

extern "C" int usl_adam_step(const usl_adam_group_t *groups, int n_groups, int64_t step, const int64_t *step_dev,
                             int zero_grad, usl_stream_t stream) {
    bool own_steps = groups != nullptr;
    for (int i = 0; groups && i < n_groups && i < USL_ADAM_MAX_GROUPS; ++i) own_steps = own_steps && groups[i].step > 0;
    if (!groups || n_groups < 1 || n_groups > USL_ADAM_MAX_GROUPS || (step < 1 && !step_dev && !own_steps)) {
        set_error("usl_adam_step: bad arguments (1..%d groups, step >= 1)", USL_ADAM_MAX_GROUPS);
        return 1;
    }
    AdamArgs A;
    const int64_t per_block = ADAM_THREADS * ADAM_VEC_PER_THREAD * 4;
    int blocks = 0;
    for (int i = 0; i < n_groups; ++i) {
        A.g[i] = groups[i];
        A.first_block[i] = blocks;
        blocks += (int)((groups[i].n + per_block - 1) / per_block);
        const double t = (double)(groups[i].step > 0 ? groups[i].step : (step < 1 ? 1 : step));
        A.step_size[i] = (float)((double)groups[i].lr / (1.0 - std::pow((double)groups[i].beta1, t)));
        A.bc2_sqrt[i] = (float)std::sqrt(1.0 - std::pow((double)groups[i].beta2, t));
    }
    A.first_block[n_groups] = blocks;
    A.n_groups = n_groups; A.zero_grad = zero_grad; A.step_dev = step_dev;
    if (blocks == 0) return 0;
    adam_kernel<<<blocks, ADAM_THREADS, 0, (cudaStream_t)stream>>>(A);
    return check_launch("usl_adam_step");
}
