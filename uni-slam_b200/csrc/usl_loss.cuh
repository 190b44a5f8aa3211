// Shared pieces of the two-phase loss: accumulator slots, ray mask, sample classes, final value.
#pragma once
#include "usl_device.cuh"

namespace usl {

#define LOSS_WARPS 8
enum { A_FS = 0, A_CENTER, A_TAIL, A_DEPTH, A_COLOR, N_FRONT, N_CENTER, N_TAIL, N_MASK, N_RAYS, N_COLOR, A_PUNC };

__device__ __forceinline__ bool ray_mask(const usl_loss_args_t &a, float gt, float punc, float depth, const float *median) {
    if (a.mode == 2) return true;                                           // "no_mask" (Mapper.py:432-440, Tracker.py:230-238)
    const bool alpha_mask = (1.0f - punc) > 0.99f;                          // Mapper.py:414-415 / Tracker.py:210-211
    if (a.mode == 0) return (gt > 0.f) && alpha_mask;                       // Mapper.py:417-420
    const float err = fabsf(gt - depth);
    return (err < 10.0f * median[0]) && alpha_mask;                         // Tracker.py:213-218
}

// sample class: 0 front, 1 center, 2 tail, 3 none (behind the surface band)
__device__ __forceinline__ int sample_class(float z, float gt, float tr, float tr04) {
    const bool front = z < (gt - tr);
    const bool back = z > (gt + tr);
    const bool center = (z > (gt - tr04)) && (z < (gt + tr04));
    if (front) return 0;
    if (center) return 1;
    if (!back) return 2;
    return 3;
}

__device__ __forceinline__ float loss_value(const usl_loss_args_t &a, const float *acc) {
    // torch.mean over an empty selection is NaN (0/0): kept
    const float fs = acc[A_FS] / acc[N_FRONT], ce = acc[A_CENTER] / acc[N_CENTER], ta = acc[A_TAIL] / acc[N_TAIL];
    const float col = acc[A_COLOR] / acc[N_COLOR], dep = acc[A_DEPTH] / acc[N_MASK];
    return a.w_sdf_fs * fs + a.w_sdf_center * ce + a.w_sdf_tail * ta + a.w_color * col + a.w_depth * dep;
}

}  // namespace usl
