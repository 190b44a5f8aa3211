// TEST INFRASTRUCTURE, NOT PRODUCT CODE: the thread function of usl_render_metrics (uni-slam_b200/csrc/usl_metrics.cuh, the
// source the CUDA kernel inlines) run for every thread of a simulated grid on the host.  See cull_host.cpp / loader.py.
#include <stdint.h>

#include "../../uni-slam_b200/csrc/usl_metrics.cuh"

// every thread of a simulated grid (nthreads = gridDim.x * 256), partial sums added in thread order
extern "C" void metrics_host(const float *gt_color, const float *gt_depth, const float *color, const float *depth, int64_t n, int64_t nthreads,
                             double *acc) {
    for (int64_t tid = 0; tid < nthreads; ++tid) {
        double se, ad, cnt;
        usl::metrics_thread(gt_color, gt_depth, color, depth, n, tid, nthreads, se, ad, cnt);
        acc[0] += se; acc[1] += ad; acc[2] += cnt;
    }
}
