// f4: marching cubes on the (slab of the) SDF volume the dense query left on the device, instead of copying
// 504 MiB to the host for skimage.measure.marching_cubes (src/utils/Mesher.py:219-258).  Three passes over a y-slab:
//
//   usl_mc_classify   per grid point: which of its three outgoing edges (+x, +y, +z) cross the level (a vertex lives on every
//                     such edge, owned by the point); per cell: its sign configuration's triangle count (case table)
//   usl_scan_u8       exclusive prefix sums of the two count arrays (block sums -> scan of block sums -> block scan + offset)
//   usl_mc_emit       vertices  p = origin + spacing * (index + t),  t = (level - v0) / (v1 - v0)  on the owned edges, and
//                     indexed triangles (a cell's edge -> owner point and axis -> that point's vertex offset + rank)
//
// Indexed output with shared vertices, numbered by owner point in the volume's memory order then by axis; faces by cell in
// the same order then by the table's order -- exactly the numbering of oracle/mc_ref.py, so results compare bit for bit.
// Volume layout: vol[(iy * nx + ix) * nz + iz] (what usl_sdf_query_grid writes), rows iy in [0, rows) of a slab; a slab
// that is not the last one carries one halo row so that the cells of its last own row are complete.
#include <cstdio>

#include "mc_tables.h"
#include "usl_device.cuh"

namespace usl {

__constant__ uint8_t c_ntri[256];
__constant__ int8_t c_tri[256][16];
static bool g_tables_loaded[16] = {false};

static int load_tables() {
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 16) { set_error("usl_mc: device ordinal %d not supported", dev); return 1; }
    if (!g_tables_loaded[dev]) {        // constant tables of this module on this device (idempotent; a race only uploads twice)
        if (cudaMemcpyToSymbol(c_ntri, USL_MC_NTRI_H, sizeof(USL_MC_NTRI_H)) != cudaSuccess ||
            cudaMemcpyToSymbol(c_tri, USL_MC_TRI_H, sizeof(USL_MC_TRI_H)) != cudaSuccess) {
            cudaGetLastError(); set_error("usl_mc: cannot upload the case tables"); return 1;
        }
        g_tables_loaded[dev] = true;
    }
    return 0;
}

struct McArgs {
    const float *vol;
    int nx, nz, rows;          // rows of this slab including the halo row (if any)
    int own_rows;              // rows whose points / cells this slab owns (rows - 1 when a halo row is present, else rows)
    float level;
    float org[3], sp[3];       // world position of grid index (0,0,0) of the WHOLE volume and the spacing
    int y_begin;               // global row index of the slab's first row
    uint8_t *pflags;           // [rows*nx*nz] bit a: the edge from this point along axis a (0 x, 1 y, 2 z) crosses the level
    uint8_t *ctri;             // [rows*nx*nz] triangles of the cell whose origin is this point (0 outside the cell range)
    const uint32_t *voff, *toff;   // exclusive prefix sums of popc(pflags) / ctri
    float *verts;              // [V,3]
    int64_t *vkeys;            // [V] global edge key = 3 * global point index + axis (nullable)
    int32_t *faces;            // [T,3] vertex indices (local to this slab's vertex array)
    int64_t ny_total;
};

__global__ void __launch_bounds__(256) mc_classify_kernel(const __grid_constant__ McArgs A) {
    const int64_t n = (int64_t)A.rows * A.nx * A.nz;
    const int64_t stride_x = A.nz, stride_y = (int64_t)A.nx * A.nz;
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (int64_t)gridDim.x * blockDim.x) {
        const int iz = (int)(p % A.nz);
        const int64_t r = p / A.nz;
        const int ix = (int)(r % A.nx), iy = (int)(r / A.nx);
        const float v = A.vol[p];
        const bool in0 = v < A.level;
        const bool hx = ix + 1 < A.nx, hy = iy + 1 < A.rows, hz = iz + 1 < A.nz;
        const bool own = iy < A.own_rows;
        uint32_t fl = 0;
        // edges along x / z of the halo row are emitted too (their owner is the next slab's first row: a merge welds them by key)
        if (hx && ((A.vol[p + stride_x] < A.level) != in0)) fl |= 1u;
        if (hy && own && ((A.vol[p + stride_y] < A.level) != in0)) fl |= 2u;
        if (hz && ((A.vol[p + 1] < A.level) != in0)) fl |= 4u;
        A.pflags[p] = (uint8_t)fl;
        uint32_t nt = 0;
        if (hx && hy && hz && own) {
            uint32_t cfg = in0 ? 1u : 0u;
            cfg |= (A.vol[p + stride_x] < A.level) ? 2u : 0u;
            cfg |= (A.vol[p + stride_y] < A.level) ? 4u : 0u;
            cfg |= (A.vol[p + stride_x + stride_y] < A.level) ? 8u : 0u;
            cfg |= (A.vol[p + 1] < A.level) ? 16u : 0u;
            cfg |= (A.vol[p + stride_x + 1] < A.level) ? 32u : 0u;
            cfg |= (A.vol[p + stride_y + 1] < A.level) ? 64u : 0u;
            cfg |= (A.vol[p + stride_x + stride_y + 1] < A.level) ? 128u : 0u;
            nt = c_ntri[cfg];
        }
        A.ctri[p] = (uint8_t)nt;
    }
}

// ---- exclusive scan of a byte array (optionally of its population counts) into uint32 ----
#define SCAN_THREADS 256
#define SCAN_ITEMS 16                                   // per thread: 4096 elements per block
__device__ __forceinline__ uint32_t scan_val(uint8_t b, int popc) { return popc ? (uint32_t)__popc((unsigned)b) : (uint32_t)b; }

__global__ void __launch_bounds__(SCAN_THREADS) scan_block_sums_kernel(const uint8_t *__restrict__ in, int64_t n, int popc,
                                                                       uint32_t *__restrict__ sums) {
    __shared__ uint32_t s_w[SCAN_THREADS / 32];
    const int64_t base = (int64_t)blockIdx.x * SCAN_THREADS * SCAN_ITEMS;
    uint32_t acc = 0;
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        const int64_t i = base + (int64_t)k * SCAN_THREADS + threadIdx.x;
        if (i < n) acc += scan_val(in[i], popc);
    }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
        for (int w = 0; w < SCAN_THREADS / 32; ++w) t += s_w[w];
        sums[blockIdx.x] = t;
    }
}

// one CTA: exclusive scan of the block sums in place; total -> total_out[0]
__global__ void __launch_bounds__(1024) scan_sums_kernel(uint32_t *__restrict__ sums, int64_t n_blocks, uint32_t *__restrict__ total_out) {
    __shared__ uint32_t s_w[32];
    __shared__ uint32_t s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int64_t base = 0; base < n_blocks; base += 1024) {
        const int64_t i = base + threadIdx.x;
        const uint32_t v = (i < n_blocks) ? sums[i] : 0u;
        uint32_t x = v;
        for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, x, o); if ((threadIdx.x & 31) >= o) x += y; }
        if ((threadIdx.x & 31) == 31) s_w[threadIdx.x >> 5] = x;
        __syncthreads();
        if (threadIdx.x < 32) {
            uint32_t w = s_w[threadIdx.x];
            for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, w, o); if (threadIdx.x >= o) w += y; }
            s_w[threadIdx.x] = w;
        }
        __syncthreads();
        const uint32_t warp_off = (threadIdx.x >> 5) ? s_w[(threadIdx.x >> 5) - 1] : 0u;
        const uint32_t carry = s_carry;
        if (i < n_blocks) sums[i] = carry + warp_off + x - v;              // exclusive
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = carry + warp_off + x;
        __syncthreads();
    }
    if (threadIdx.x == 0) total_out[0] = s_carry;
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_apply_kernel(const uint8_t *__restrict__ in, int64_t n, int popc,
                                                                  const uint32_t *__restrict__ sums, uint32_t *__restrict__ out) {
    // element order inside a block: thread t owns the SCAN_ITEMS consecutive elements base + t*SCAN_ITEMS .. (blocked arrangement)
    __shared__ uint32_t s_w[SCAN_THREADS / 32];
    const int64_t base = (int64_t)blockIdx.x * SCAN_THREADS * SCAN_ITEMS + (int64_t)threadIdx.x * SCAN_ITEMS;
    uint32_t v[SCAN_ITEMS];
    uint32_t acc = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        v[k] = (base + k < n) ? scan_val(in[base + k], popc) : 0u;
        acc += v[k];
    }
    uint32_t x = acc;
    for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, x, o); if ((threadIdx.x & 31) >= o) x += y; }
    if ((threadIdx.x & 31) == 31) s_w[threadIdx.x >> 5] = x;
    __syncthreads();
    uint32_t woff = 0;
    for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) woff += s_w[w];
    uint32_t run = sums[blockIdx.x] + woff + x - acc;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        if (base + k < n) out[base + k] = run;
        run += v[k];
    }
}

__global__ void __launch_bounds__(256) mc_emit_kernel(const __grid_constant__ McArgs A) {
    const int64_t n = (int64_t)A.rows * A.nx * A.nz;
    const int64_t stride_x = A.nz, stride_y = (int64_t)A.nx * A.nz;
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t fl = A.pflags[p];
        const uint32_t nt = A.ctri[p];
        if (!fl && !nt) continue;
        const int iz = (int)(p % A.nz);
        const int64_t r = p / A.nz;
        const int ix = (int)(r % A.nx), iy = (int)(r / A.nx);
        if (fl) {
            const float v0 = A.vol[p];
            uint32_t vo = A.voff[p];
            const float idx[3] = {(float)ix, (float)(iy + A.y_begin), (float)iz};
            const int64_t gp = ((int64_t)(iy + A.y_begin) * A.nx + ix) * A.nz + iz;
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                if (!((fl >> a) & 1u)) continue;
                const float v1 = A.vol[p + (a == 0 ? stride_x : (a == 1 ? stride_y : 1))];
                const float t = __fdiv_rn(__fsub_rn(A.level, v0), __fsub_rn(v1, v0));
#pragma unroll
                for (int d = 0; d < 3; ++d) {
                    const float c = (d == a) ? __fadd_rn(idx[d], t) : idx[d];
                    A.verts[(int64_t)vo * 3 + d] = __fadd_rn(A.org[d], __fmul_rn(A.sp[d], c));
                }
                if (A.vkeys) A.vkeys[vo] = gp * 3 + a;
                ++vo;
            }
        }
        if (nt) {
            uint32_t cfg = (A.vol[p] < A.level) ? 1u : 0u;
            cfg |= (A.vol[p + stride_x] < A.level) ? 2u : 0u;
            cfg |= (A.vol[p + stride_y] < A.level) ? 4u : 0u;
            cfg |= (A.vol[p + stride_x + stride_y] < A.level) ? 8u : 0u;
            cfg |= (A.vol[p + 1] < A.level) ? 16u : 0u;
            cfg |= (A.vol[p + stride_x + 1] < A.level) ? 32u : 0u;
            cfg |= (A.vol[p + stride_y + 1] < A.level) ? 64u : 0u;
            cfg |= (A.vol[p + stride_x + stride_y + 1] < A.level) ? 128u : 0u;
            int32_t *f = A.faces + (int64_t)A.toff[p] * 3;
            for (uint32_t k = 0; k < nt * 3; ++k) {
                const int e = c_tri[cfg][k];
                const int a = e >> 2, bu = e & 1, bv = (e >> 1) & 1;
                // owner point of edge e: offsets of the two non-axis coordinates, in the order (y,z) / (x,z) / (x,y)
                const int dx = (a == 0) ? 0 : bu;
                const int dy = (a == 0) ? bu : ((a == 1) ? 0 : bv);
                const int dz = (a == 2) ? 0 : bv;
                const int64_t q = p + dx * stride_x + dy * stride_y + dz;
                const uint32_t qf = A.pflags[q];
                f[k] = (int32_t)(A.voff[q] + __popc(qf & ((1u << a) - 1u)));
            }
        }
    }
}

}  // namespace usl

using namespace usl;

extern "C" {

int usl_scan_u8(const uint8_t *in, int64_t n, int popcount, uint32_t *out, uint32_t *block_sums, uint32_t *total, usl_stream_t stream) {
    if (!in || !out || !block_sums || !total || n < 0) { set_error("usl_scan_u8: bad arguments"); return 1; }
    cudaStream_t s = (cudaStream_t)stream;
    const int64_t per_block = (int64_t)SCAN_THREADS * SCAN_ITEMS;
    const int64_t nb = (n + per_block - 1) / per_block;
    if (nb == 0) { cudaMemsetAsync(total, 0, sizeof(uint32_t), s); return 0; }
    scan_block_sums_kernel<<<(unsigned)nb, SCAN_THREADS, 0, s>>>(in, n, popcount, block_sums);
    scan_sums_kernel<<<1, 1024, 0, s>>>(block_sums, nb, total);
    scan_apply_kernel<<<(unsigned)nb, SCAN_THREADS, 0, s>>>(in, n, popcount, block_sums, out);
    return check_launch("usl_scan_u8");
}

int usl_scan_u8_blocks(int64_t n, int64_t *n_blocks) {
    if (!n_blocks || n < 0) { set_error("usl_scan_u8_blocks: bad arguments"); return 1; }
    *n_blocks = (n + (int64_t)SCAN_THREADS * SCAN_ITEMS - 1) / ((int64_t)SCAN_THREADS * SCAN_ITEMS);
    return 0;
}

static int fill_mc(McArgs &A, const usl_mc_args_t *a, const char *who) {
    if (!a || !a->vol || a->nx < 2 || a->nz < 2 || a->rows < 1 || a->own_rows < 0 || a->own_rows > a->rows || !a->pflags || !a->ctri) {
        set_error("%s: bad arguments", who); return 1;
    }
    A.vol = a->vol; A.nx = a->nx; A.nz = a->nz; A.rows = a->rows; A.own_rows = a->own_rows; A.level = a->level; A.y_begin = a->y_begin;
    for (int d = 0; d < 3; ++d) { A.org[d] = a->origin[d]; A.sp[d] = a->spacing[d]; }
    A.pflags = a->pflags; A.ctri = a->ctri; A.voff = a->voff; A.toff = a->toff; A.verts = a->verts; A.vkeys = a->vkeys; A.faces = a->faces;
    A.ny_total = 0;
    return load_tables();
}

static unsigned mc_grid(int64_t n) {
    int dev = 0, n_sm = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    int64_t b = (n + 255) / 256;
    const int64_t cap = (int64_t)(n_sm > 0 ? n_sm : 1) * 16;
    return (unsigned)(b < cap ? (b < 1 ? 1 : b) : cap);
}

int usl_mc_classify(const usl_mc_args_t *a, usl_stream_t stream) {
    McArgs A;
    if (fill_mc(A, a, "usl_mc_classify")) return 1;
    const int64_t n = (int64_t)A.rows * A.nx * A.nz;
    mc_classify_kernel<<<mc_grid(n), 256, 0, (cudaStream_t)stream>>>(A);
    return check_launch("usl_mc_classify");
}

int usl_mc_emit(const usl_mc_args_t *a, usl_stream_t stream) {
    McArgs A;
    if (fill_mc(A, a, "usl_mc_emit")) return 1;
    if (!A.voff || !A.toff || !A.verts || !A.faces) { set_error("usl_mc_emit: offsets and output buffers are required"); return 1; }
    const int64_t n = (int64_t)A.rows * A.nx * A.nz;
    mc_emit_kernel<<<mc_grid(n), 256, 0, (cudaStream_t)stream>>>(A);
    return check_launch("usl_mc_emit");
}

}  // extern "C"
