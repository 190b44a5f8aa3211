// TEST INFRASTRUCTURE, NOT PRODUCT CODE: the element function of usl_render_metrics (uni-slam_b200/csrc/usl_metrics.cuh, the
// source the CUDA kernel inlines) run serially over a frame on the host.  See cull_host.cpp / loader.py.
#include <stdint.h>

#include "../../uni-slam_b200/csrc/usl_metrics.cuh"

extern "C" void metrics_host(const float *gt_color, const float *gt_depth, const float *color, const float *depth, int64_t n, double *acc) {
    for (int64_t i = 0; i < n; ++i) {
        double s, a;
        if (usl::metrics_pixel(gt_color + i * 3, color + i * 3, gt_depth[i], depth[i], s, a)) { acc[0] += s; acc[1] += a; acc[2] += 1.0; }
    }
}
