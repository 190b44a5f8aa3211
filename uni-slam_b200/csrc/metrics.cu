// f3, the consumer of the whole-frame renderer at the end of a run: eval_rendering (src/tools/eval_recon.py:235-307)
// renders every 5th frame and computes, per frame, PSNR over the pixels with sensor depth and the depth L1 -- in the
// reference two boolean-index compactions (each a host sync), an mse_loss, a log10 and an .item().  Here: one streaming
// reduction per frame over the renderer's output buffers where they lie (816 000 pixels x 32 B = 26 MB: a few microseconds at
// HBM rate, launch-latency bound), sums kept in double and left on the device, so a sequence is evaluated without a host
// read until its end.  (MS-SSIM and LPIPS are third-party networks and stay host code.)
#include "usl_device.cuh"
#include "usl_metrics.cuh"

namespace usl {

constexpr int METRICS_THREADS = 256;

__global__ void __launch_bounds__(METRICS_THREADS) render_metrics_kernel(const float *__restrict__ gt_color, const float *__restrict__ gt_depth,
                                                                         const float *__restrict__ color, const float *__restrict__ depth,
                                                                         int64_t n, double *__restrict__ acc) {
    double se, ad, cnt;
    metrics_thread(gt_color, gt_depth, color, depth, n, (int64_t)blockIdx.x * METRICS_THREADS + threadIdx.x, (int64_t)gridDim.x * METRICS_THREADS, se, ad, cnt);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        se += __shfl_down_sync(0xffffffffu, se, o);
        ad += __shfl_down_sync(0xffffffffu, ad, o);
        cnt += __shfl_down_sync(0xffffffffu, cnt, o);
    }
    __shared__ double s_part[3][METRICS_THREADS / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) { s_part[0][warp] = se; s_part[1][warp] = ad; s_part[2][warp] = cnt; }
    __syncthreads();
    if (threadIdx.x < 3) {
        double t = 0.0;
        for (int w = 0; w < METRICS_THREADS / 32; ++w) t += s_part[threadIdx.x][w];
        if (t != 0.0) atomicAdd(acc + threadIdx.x, t);
    }
}

}  // namespace usl

using namespace usl;

extern "C" {

int usl_render_metrics(const float *gt_color, const float *gt_depth, const float *color, const float *depth, int64_t n, double *acc,
                       usl_stream_t stream) {
    if (n <= 0) return 0;
    if (!gt_color || !gt_depth || !color || !depth || !acc) { set_error("usl_render_metrics: null argument"); return 1; }
    int dev = 0, n_sm = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    int64_t blocks = (n + METRICS_THREADS - 1) / METRICS_THREADS;
    const int64_t cap = (int64_t)(n_sm > 0 ? n_sm : 1) * 4;          // few, long-running CTAs: 3 double atomics per CTA
    if (blocks > cap) blocks = cap;
    render_metrics_kernel<<<(unsigned)blocks, METRICS_THREADS, 0, (cudaStream_t)stream>>>(gt_color, gt_depth, color, depth, n, acc);
    return check_launch("usl_render_metrics");
}

}  // extern "C"
