// Stand-alone multi-resolution hash-grid encoding kernels: the tinycudann.Encoding seam (B1).
// Replaces tcnn kernel_grid / kernel_grid_backward / kernel_grid_backward_input as used at
// src/networks/decoders.py:101-103 and by loss.backward() (src/Mapper.py:444, src/Tracker.py:241).
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>

#include "usl_device.cuh"

namespace usl {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int check_launch(const char *what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(e));
        return 1;
    }
    return 0;
}

// thread per (point, level): blockIdx.y = level keeps a block inside one level's table.
__global__ void __launch_bounds__(256) encode_fwd_kernel(usl_grid_t g, const float2 *__restrict__ table,
                                                         const float *__restrict__ x, int64_t n,
                                                         float2 *__restrict__ y) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int l = blockIdx.y;
    const float x0 = x[i * 3 + 0], x1 = x[i * 3 + 1], x2 = x[i * 3 + 2];
    float2 f, df[3];
    level_interp<false, true>(g.levels[l], table, x0, x1, x2, f, df);
    y[i * g.n_levels + l] = f;
}

__global__ void __launch_bounds__(256) encode_bwd_params_kernel(usl_grid_t g, const float *__restrict__ x,
                                                                const float2 *__restrict__ dy, int64_t n,
                                                                float2 *__restrict__ grad) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int l = blockIdx.y;
    const usl_level_t &lv = g.levels[l];
    const Cell c = make_cell(lv, x[i * 3 + 0], x[i * 3 + 1], x[i * 3 + 2]);
    uint32_t idx[8];
    float wt[8];
    corner_indices(lv, c, idx);
    corner_weights(c, wt);
    const float2 d = dy[i * g.n_levels + l];
    scatter_level(grad + lv.offset, idx, wt, d.x, d.y);
}

__global__ void __launch_bounds__(256) encode_bwd_input_kernel(usl_grid_t g, const float2 *__restrict__ table,
                                                               const float *__restrict__ x,
                                                               const float2 *__restrict__ dy, int64_t n,
                                                               float *__restrict__ dx) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float x0 = x[i * 3 + 0], x1 = x[i * 3 + 1], x2 = x[i * 3 + 2];
    float a0 = 0.f, a1 = 0.f, a2 = 0.f;
    for (int l = 0; l < g.n_levels; ++l) {
        float2 f, df[3];
        level_interp<true, true>(g.levels[l], table, x0, x1, x2, f, df);
        const float2 d = dy[i * g.n_levels + l];
        a0 += d.x * df[0].x + d.y * df[0].y;
        a1 += d.x * df[1].x + d.y * df[1].y;
        a2 += d.x * df[2].x + d.y * df[2].y;
    }
    dx[i * 3 + 0] = a0; dx[i * 3 + 1] = a1; dx[i * 3 + 2] = a2;
}

__global__ void __launch_bounds__(256) corner_indices_kernel(usl_grid_t g, const float *__restrict__ x, int64_t n,
                                                             uint32_t *__restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int l = blockIdx.y;
    const Cell c = make_cell(g.levels[l], x[i * 3 + 0], x[i * 3 + 1], x[i * 3 + 2]);
    uint32_t idx[8];
    corner_indices(g.levels[l], c, idx);
#pragma unroll
    for (int k = 0; k < 8; ++k) out[(i * g.n_levels + l) * 8 + k] = idx[k];
}

// ---- measurement utilities: L2-resident random 8-byte gather / vector-atomic scatter ceilings ----
__device__ __forceinline__ uint32_t mix32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}

__global__ void __launch_bounds__(256) bench_gather_kernel(const float2 *__restrict__ table, uint32_t entries, int64_t n,
                                                           int per_thread, float *__restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float ax = 0.f, ay = 0.f;
    uint32_t h = mix32((uint32_t)i * 2654435761u + 12345u);
    for (int k = 0; k < per_thread; k += 8) {
        float2 v[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) { h = mix32(h + q); v[q] = __ldg(table + (h % entries)); }
#pragma unroll
        for (int q = 0; q < 8; ++q) { ax += v[q].x; ay += v[q].y; }
    }
    if (ax == 123.456f) out[i] = ax + ay;            // keep the loads alive
}

// Coalesced 16-byte reads of a buffer, `repeats` passes: with a buffer that fits L2 this measures the L2 -> SM read
// bandwidth (the first pass warms it), with a larger one the HBM read bandwidth.
__global__ void __launch_bounds__(256) bench_stream_read_kernel(const float4 *__restrict__ buf, int64_t n4, int repeats,
                                                                float *__restrict__ out) {
    float acc = 0.f;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int r = 0; r < repeats; ++r) {
        for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i + 3 * stride < n4; i += 4 * stride) {
            const float4 a = __ldcg(buf + i), b = __ldcg(buf + i + stride), c = __ldcg(buf + i + 2 * stride), d = __ldcg(buf + i + 3 * stride);
            acc += (a.x + a.y + a.z + a.w) + (b.x + b.y + b.z + b.w) + (c.x + c.y + c.z + c.w) + (d.x + d.y + d.z + d.w);
        }
    }
    if (acc == 123.456f) out[0] = acc;                   // keep the loads alive
}

// mode 0: every lane its own random entry; 1: lane pairs share a 16 B slot; 3: lane quads share a 32 B sector;
// 4: a warp covers 32 consecutive entries; 2: one float4 atomic per lane at a random 16 B slot
__global__ void __launch_bounds__(256) bench_scatter_kernel(float2 *__restrict__ table, uint32_t entries, int64_t n,
                                                            int per_thread, int mode) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t lane = threadIdx.x & 31;
    // 5: all 32 lanes the SAME entry; 6: runs of 4 lanes the same entry; 7: runs of 2 lanes the same entry
    const uint32_t grp = (mode == 1 || mode == 7) ? (uint32_t)(i >> 1) : (mode == 3 || mode == 6) ? (uint32_t)(i >> 2) : (mode == 4 || mode == 5) ? (uint32_t)(i >> 5) : (uint32_t)i;
    uint32_t h = mix32(grp * 2654435761u + 12345u);
    for (int k = 0; k < per_thread; ++k) {
        h = mix32(h + k);
        uint32_t e = h % entries;
        if (mode == 1) e = (e & ~1u) | (lane & 1u);
        else if (mode == 3) e = (e & ~3u) | (lane & 3u);
        else if (mode == 4) e = (e & ~31u) | lane;
        if (mode == 2) atomicAdd(reinterpret_cast<float4 *>(table + (e & ~1u)), make_float4(1.0f, 0.5f, 0.25f, 2.0f));
        else atomicAdd(table + e, make_float2(1.0f, 0.5f));
    }
}

static int check_grid(const usl_grid_t *g) {
    if (!g || g->n_levels <= 0 || g->n_levels > USL_MAX_LEVELS) {
        set_error("invalid usl_grid_t (n_levels out of range)");
        return 1;
    }
    return 0;
}

}  // namespace usl

using namespace usl;

extern "C" {

const char *usl_last_error(void) { return g_err; }
int usl_version(void) { return 100; }

// tcnn GridEncoding constructor arithmetic, host libm in fp32.
int usl_grid_build(int n_levels, int log2_hashmap_size, int base_resolution, double per_level_scale,
                   usl_grid_t *out) {
    if (!out || n_levels <= 0 || n_levels > USL_MAX_LEVELS || log2_hashmap_size < 3 || log2_hashmap_size > 30) {
        set_error("usl_grid_build: bad arguments");
        return 1;
    }
    memset(out, 0, sizeof(*out));
    const float log2_pls = log2f((float)per_level_scale);
    uint32_t offset = 0;
    for (int l = 0; l < n_levels; ++l) {
        const float scale = exp2f((float)l * log2_pls) * (float)base_resolution - 1.0f;
        const uint32_t res = (uint32_t)ceilf(scale) + 1u;
        const double dense = (double)res * res * res;
        const uint32_t max_params = 0xFFFFFFFFu / 2u;
        uint32_t cnt = dense > (double)max_params ? max_params : (uint32_t)dense;
        cnt = (cnt + 7u) / 8u * 8u;
        const uint32_t cap = 1u << log2_hashmap_size;
        if (cnt > cap) cnt = cap;
        out->levels[l].scale = scale;
        out->levels[l].res = res;
        out->levels[l].size = cnt;
        out->levels[l].offset = offset;
        out->levels[l].hashed = dense > (double)cnt ? 1u : 0u;
        offset += cnt;
    }
    out->n_levels = n_levels;
    out->total_entries = offset;
    return 0;
}

int usl_grid_encode_fwd(const usl_grid_t *g, const float *params, const float *x, int64_t n, float *y,
                        usl_stream_t stream) {
    if (check_grid(g)) return 1;
    if (n <= 0) return 0;
    dim3 grid((unsigned)((n + 255) / 256), g->n_levels);
    encode_fwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(*g, (const float2 *)params, x, n, (float2 *)y);
    return check_launch("usl_grid_encode_fwd");
}

int usl_grid_encode_bwd_params(const usl_grid_t *g, const float *x, const float *dy, int64_t n,
                               float *grad_params, usl_stream_t stream) {
    if (check_grid(g)) return 1;
    if (n <= 0) return 0;
    dim3 grid((unsigned)((n + 255) / 256), g->n_levels);
    encode_bwd_params_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(*g, x, (const float2 *)dy, n, (float2 *)grad_params);
    return check_launch("usl_grid_encode_bwd_params");
}

int usl_grid_encode_bwd_input(const usl_grid_t *g, const float *params, const float *x, const float *dy,
                              int64_t n, float *dx, usl_stream_t stream) {
    if (check_grid(g)) return 1;
    if (n <= 0) return 0;
    encode_bwd_input_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        *g, (const float2 *)params, x, (const float2 *)dy, n, dx);
    return check_launch("usl_grid_encode_bwd_input");
}

int usl_bench_stream_read(const float *buf, int64_t n_floats, int repeats, float *out, usl_stream_t stream) {
    if (!buf || n_floats < 4 || repeats < 1) { set_error("usl_bench_stream_read: bad arguments"); return 1; }
    int dev = 0, n_sm = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    bench_stream_read_kernel<<<(unsigned)(n_sm * 8), 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float4 *>(buf), n_floats / 4, repeats, out);
    return check_launch("usl_bench_stream_read");
}

int usl_bench_gather(const float *table, uint32_t entries, int64_t n_threads, int per_thread, float *out,
                     usl_stream_t stream) {
    bench_gather_kernel<<<(unsigned)((n_threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>((const float2 *)table, entries, n_threads, per_thread, out);
    return check_launch("usl_bench_gather");
}

int usl_bench_scatter(float *table, uint32_t entries, int64_t n_threads, int per_thread, int mode, usl_stream_t stream) {
    bench_scatter_kernel<<<(unsigned)((n_threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>((float2 *)table, entries, n_threads, per_thread, mode);
    return check_launch("usl_bench_scatter");
}

int usl_grid_corner_indices(const usl_grid_t *g, const float *x, int64_t n, uint32_t *idx, usl_stream_t stream) {
    if (check_grid(g)) return 1;
    if (n <= 0) return 0;
    dim3 grid((unsigned)((n + 255) / 256), g->n_levels);
    corner_indices_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(*g, x, n, idx);
    return check_launch("usl_grid_corner_indices");
}

}  // extern "C"
