// Element and thread functions of usl_render_metrics (metrics.cu): eval_rendering's per-pixel terms (src/tools/eval_recon.py:276-286)
// and a thread's grid-stride partial sums.
// Like usl_cull.cuh it also compiles with plain g++ for the host-side test harness (tests/host_harness); the library never
// runs it on the CPU.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define USLM_HD __host__ __device__ __forceinline__
#else
#define USLM_HD static inline
#endif

namespace usl {

// A pixel counts when the sensor depth is positive (gt_depth > 0).  se = sum over the three channels of (gt - rendered)^2,
// ad = |gt_depth - depth|, both in double: the reference's dataset colour is float64 and render_img returns depth as float64,
// so its differences are taken in double as well.
USLM_HD bool metrics_pixel(const float *gt_color, const float *color, float gt_depth, float depth, double &se, double &ad) {
    if (!(gt_depth > 0.f)) return false;
    double s = 0.0;
    for (int c = 0; c < 3; ++c) {
        const double e = (double)gt_color[c] - (double)color[c];
        s += e * e;
    }
    se = s;
    ad = fabs((double)gt_depth - (double)depth);
    return true;
}

// One thread of render_metrics_kernel: its grid-stride partial sums (thread `tid` of `nthreads`).
USLM_HD void metrics_thread(const float *gt_color, const float *gt_depth, const float *color, const float *depth, int64_t n, int64_t tid,
                            int64_t nthreads, double &se, double &ad, double &cnt) {
    se = 0.0; ad = 0.0; cnt = 0.0;
    for (int64_t i = tid; i < n; i += nthreads) {
        double s, a;
        if (metrics_pixel(gt_color + i * 3, color + i * 3, gt_depth[i], depth[i], s, a)) { se += s; ad += a; cnt += 1.0; }
    }
}

}  // namespace usl
