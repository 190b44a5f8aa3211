// field_fwd with the tangent contraction on the 5th-generation tensor cores (tcgen05 + TMEM).
//
// When ray gradients are needed (tracker pose, mapper joint-opt) the forward pass pushes three tangent vectors
// (d f / d x_0..2, 32 features each) through the first decoder layer: per point a [3 x 32] . [32 x 16] product,
// 1536 FMAs = 57 % of the CUDA-core FMA-pipe work of field_fwd_kernel<JAC>.  Here a CTA of 128 threads owns 128
// points = the 128 rows (TMEM lanes) of three M=128, N=16 accumulators; every 4 levels (8 features = one
// K-step of kind::tf32) each thread writes its three tangent rows into a K-major, un-swizzled shared-memory
// operand tile and one thread issues 3 tcgen05.mma; tcgen05.ld (32x32b) returns row t to thread t at the end.
// The VALUE path (h = W1 f + b1) stays in fp32 on the CUDA cores: outputs keep the 1e-4 parity bar, while the
// tangents only feed gradients (1e-3 bar) where TF32 operands (round-to-nearest, ~2.4e-4 per operand) are ample.
#include "usl_async.cuh"
#include "usl_field.cuh"

namespace usl {

#define TC_THREADS 128      // compute threads = points = accumulator rows per CTA
#define TC_CTA_THREADS 160  // + one MMA-issuer warp (warp 4): compute warps never block on a CTA-wide barrier
#define TC_TMEM_COLS 64      // 3 accumulators x 16 fp32 columns, rounded up to a power of two >= 32

struct TcSmem {
    MlpSmem mlp;
    alignas(128) uint32_t b[4][2][2][8][4];        // W1 as tf32: [chunk][n-group][k-chunk][n & 7][k & 3]      2 KB
    alignas(128) uint32_t a[2][3][16][2][8][4];    // tangents:   [buf][dim][row-group][k-chunk][row & 7][k & 3] 24 KB
    alignas(8) uint64_t full[2];                   // operand buffer written by all 128 compute threads
    alignas(8) uint64_t empty[2];                  // operand buffer consumed by the tensor core (tcgen05.commit)
    alignas(8) uint64_t done;                      // all accumulators complete
    uint32_t tmem_base;
};

// UMMA shared-memory matrix descriptor, K-major, SWIZZLE_NONE: core matrices of 8 rows x 16 bytes;
// LBO = byte distance between the two core matrices along K, SBO = between 8-row groups along M/N.
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
// instruction descriptor: D fp32, A/B tf32, both K-major, N = 16, M = 128
#define TC_IDESC ((1u << 4) | (2u << 7) | (2u << 10) | ((16u >> 3) << 17) | ((128u >> 4) << 24))

__device__ __forceinline__ uint32_t to_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t accumulate) {
    asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}"
                 ::"r"(tmem_d), "l"(a_desc), "l"(b_desc), "r"(TC_IDESC), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float v[16]) {
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                   "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr) : "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

struct FieldTcArgs {
    usl_field_t f;
    usl_points_t p;
    float *raw, *feat, *jac;
};

// same point loader as field.cu (kept local: this translation unit is self-contained)
__device__ __forceinline__ bool tc_load_point(const usl_points_t &p, const usl_field_t &f, int64_t i, float xc[3], float gate[3]) {
    float x[3];
    if (p.x) {
#pragma unroll
        for (int d = 0; d < 3; ++d) x[d] = p.x[i * 3 + d];
    } else {
        const int64_t r = i / p.S;
        if (p.valid && !p.valid[r]) return false;
        const float z = p.z[i];
#pragma unroll
        for (int d = 0; d < 3; ++d) x[d] = norm_coord(p.rays_o[r * 3 + d], p.rays_d[r * 3 + d], z, f.bound_lo[d], f.bound_hi[d]);
    }
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        xc[d] = fminf(fmaxf(x[d], 0.f), 1.f);
        gate[d] = (x[d] >= 0.f && x[d] <= 1.f) ? 1.f : 0.f;
    }
    return true;
}

template <bool SAVE_FEAT>
#ifndef USL_TC_MINB
#define USL_TC_MINB 5
#endif
__global__ void __launch_bounds__(TC_CTA_THREADS, USL_TC_MINB) field_fwd_tc_kernel(const __grid_constant__ FieldTcArgs A) {
    __shared__ TcSmem S;
    const int gi = blockIdx.y;
    const int tid = threadIdx.x, warp = tid >> 5;
    const usl_mlp_t &m = A.f.mlp[gi];
    stage_mlp(m, S.mlp);
    // W1[j][k] (row-major, = N x K K-major) -> tf32, UMMA canonical layout per K-chunk of 8
    for (int e = tid; e < USL_HID * USL_IN; e += TC_CTA_THREADS) {
        const int j = e / USL_IN, k = e % USL_IN;
        S.b[k >> 3][j >> 3][(k & 7) >> 2][j & 7][k & 3] = to_tf32(m.w1[e]);
    }
    if (tid == 0) {
        mbar_init(&S.full[0], TC_THREADS); mbar_init(&S.full[1], TC_THREADS);
        mbar_init(&S.empty[0], 1); mbar_init(&S.empty[1], 1); mbar_init(&S.done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 4) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&S.tmem_base)), "r"(TC_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = S.tmem_base;

    if (warp == 4) {
        // ---- MMA issuer: waits for an operand buffer, issues 3 tcgen05.mma, commits completion to mbarriers ----
        if (tid == TC_THREADS) {
#pragma unroll 1
            for (int c = 0; c < 4; ++c) {
                const int buf = c & 1;
                mbar_wait(&S.full[buf], (uint32_t)(c >> 1) & 1u);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint64_t bd = umma_desc(smem_u32(&S.b[c]), 128, 256);
#pragma unroll
                for (int d = 0; d < 3; ++d) umma_tf32(tmem + d * 16, umma_desc(smem_u32(&S.a[buf][d]), 128, 256), bd, c > 0 ? 1u : 0u);
                umma_commit(&S.empty[buf]);
                if (c == 3) umma_commit(&S.done);
            }
        }
        __syncwarp();
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();                                   // matches the compute warps' final barrier
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TC_TMEM_COLS) : "memory");
        return;
    }

    const int64_t n = A.p.n;
    const int64_t i = (int64_t)blockIdx.x * TC_THREADS + tid;
    float xc[3] = {0.f, 0.f, 0.f}, gate[3] = {0.f, 0.f, 0.f};
    const bool active = (i < n) && tc_load_point(A.p, A.f, i, xc, gate);
    const usl_grid_t &g = A.f.grid[gi];
    const float2 *table = reinterpret_cast<const float2 *>(A.f.table[gi]);
    float2 *fo = SAVE_FEAT ? reinterpret_cast<float2 *>(A.feat) + ((int64_t)gi * g.n_levels) * n + i : nullptr;
    if (SAVE_FEAT && i < n && gi == 0) {                   // clamped coordinates for the backward pass (x0 = -1: filtered point)
        float *xs = A.feat + (int64_t)2 * (USL_IN + USL_HID) * n + i;
        __stcs(xs, active ? xc[0] : -1.0f);
        __stcs(xs + n, xc[1]);
        __stcs(xs + 2 * n, xc[2]);
    }

    float h[USL_HID];
#pragma unroll
    for (int j = 0; j < USL_HID; ++j) h[j] = S.mlp.b1[j];
    const int rg = tid >> 3, rr = tid & 7;

#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
        const int buf = c & 1;
        if (c >= 2) mbar_wait(&S.empty[buf], 0);           // the MMAs of chunk c-2 have consumed this operand buffer
#ifndef USL_TC_LEVEL_UNROLL
#define USL_TC_LEVEL_UNROLL 4
#endif
        constexpr int kLevelUnroll = USL_TC_LEVEL_UNROLL;
#pragma unroll kLevelUnroll
        for (int q = 0; q < 4; ++q) {
            const int l = 4 * c + q;
            float2 f = make_float2(0.f, 0.f), df[3] = {f, f, f};
            if (active) level_interp<true>(g.levels[l], table, xc[0], xc[1], xc[2], f, df);
            if (SAVE_FEAT && active) __stcs(fo + (int64_t)l * n, f);
            const float4 *wa = reinterpret_cast<const float4 *>(S.mlp.w1t[2 * l]);
            const float4 *wb = reinterpret_cast<const float4 *>(S.mlp.w1t[2 * l + 1]);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float4 a = wa[e], b = wb[e];
                h[4 * e + 0] = fmaf(b.x, f.y, fmaf(a.x, f.x, h[4 * e + 0]));
                h[4 * e + 1] = fmaf(b.y, f.y, fmaf(a.y, f.x, h[4 * e + 1]));
                h[4 * e + 2] = fmaf(b.z, f.y, fmaf(a.z, f.x, h[4 * e + 2]));
                h[4 * e + 3] = fmaf(b.w, f.y, fmaf(a.w, f.x, h[4 * e + 3]));
            }
            // this level's two K columns (k = 2q, 2q+1 of the chunk) go straight to the operand tile: no tangent registers
#pragma unroll
            for (int d = 0; d < 3; ++d)
                *reinterpret_cast<uint2 *>(&S.a[buf][d][rg][q >> 1][rr][2 * (q & 1)]) = make_uint2(to_tf32(df[d].x), to_tf32(df[d].y));
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy stores -> visible to the tensor core
        mbar_arrive(&S.full[buf]);
    }
    if (SAVE_FEAT && active) {
        float *ho = A.feat + (int64_t)2 * USL_IN * n + ((int64_t)gi * USL_HID) * n + i;
#pragma unroll
        for (int j = 0; j < USL_HID; ++j) __stcs(ho + (int64_t)j * n, h[j]);
    }
    mbar_wait(&S.done, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    // ---- tail: value path first, then ONE tangent direction at a time (16 TMEM columns -> 16 registers) so the
    //      accumulators never occupy 48 registers at once ----
    uint32_t m1 = 0, m2 = 0;
#pragma unroll
    for (int j = 0; j < USL_HID; ++j) { if (h[j] > 0.f) m1 |= 1u << j; h[j] = fmaxf(h[j], 0.f); }
    float out[4], tout[4][3], da[4];
    {
        float u[4] = {S.mlp.bo[0], S.mlp.bo[1], S.mlp.bo[2], S.mlp.bo[3]};
        if (m.n_hidden == 2) {
            for (int q = 0; q < USL_HID; ++q) {
                float sacc = S.mlp.b2[q];
#pragma unroll
                for (int j = 0; j < USL_HID; ++j) sacc = fmaf(S.mlp.w2[q][j], h[j], sacc);
                if (sacc > 0.f) {
                    m2 |= 1u << q;
#pragma unroll
                    for (int o = 0; o < 4; ++o) u[o] = fmaf(S.mlp.wo[o][q], sacc, u[o]);
                }
            }
        } else {
#pragma unroll
            for (int o = 0; o < 4; ++o)
#pragma unroll
                for (int j = 0; j < USL_HID; ++j) u[o] = fmaf(S.mlp.wo[o][j], h[j], u[o]);
        }
#pragma unroll
        for (int o = 0; o < 4; ++o) { out[o] = act_fwd(m.out_act, u[o]); da[o] = act_bwd(m.out_act, out[o]); }
    }
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        float t1[USL_HID];
        tmem_ld16(tmem + lane_base + d * 16, t1);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int j = 0; j < USL_HID; ++j) t1[j] = ((m1 >> j) & 1u) ? t1[j] : 0.f;
        float tu[4] = {0.f, 0.f, 0.f, 0.f};
        if (m.n_hidden == 2) {
            for (int q = 0; q < USL_HID; ++q) {
                if (!((m2 >> q) & 1u)) continue;
                float ts = 0.f;
#pragma unroll
                for (int j = 0; j < USL_HID; ++j) ts = fmaf(S.mlp.w2[q][j], t1[j], ts);
#pragma unroll
                for (int o = 0; o < 4; ++o) tu[o] = fmaf(S.mlp.wo[o][q], ts, tu[o]);
            }
        } else {
#pragma unroll
            for (int o = 0; o < 4; ++o)
#pragma unroll
                for (int j = 0; j < USL_HID; ++j) tu[o] = fmaf(S.mlp.wo[o][j], t1[j], tu[o]);
        }
#pragma unroll
        for (int o = 0; o < 4; ++o) tout[o][d] = da[o] * tu[o];
    }
    if (active) {
        if (gi == 0) {
            A.raw[i * 4 + 3] = out[0];
#pragma unroll
            for (int d = 0; d < 3; ++d) __stcs(A.jac + (int64_t)(9 + d) * n + i, tout[0][d] * gate[d]);   // component-major: coalesced
        } else {
#pragma unroll
            for (int o = 0; o < 3; ++o) {
                A.raw[i * 4 + o] = out[o];
#pragma unroll
                for (int d = 0; d < 3; ++d) __stcs(A.jac + (int64_t)(o * 3 + d) * n + i, tout[o][d] * gate[d]);
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();                                       // the issuer warp deallocates TMEM after this barrier
}

}  // namespace usl

using namespace usl;

// The tensor-core variant of usl_field_fwd with a Jacobian (same arguments, same outputs): an explicit entry point, so a caller
// chooses it by name, not through process-wide state.
extern "C" int usl_field_fwd_tc(const usl_field_t *f, const usl_points_t *p, float *raw, float *feat, float *jac, usl_stream_t stream) {
    if (!f || !p || !raw || !jac) { set_error("usl_field_fwd_tc: field, points, raw and jac are required"); return 1; }
    if (p->n <= 0) return 0;
    if (p->sample_major) { set_error("usl_field_fwd_tc: ray-major point order only"); return 1; }
    for (int gi = 0; gi < 2; ++gi)
        if (f->grid[gi].n_levels != USL_IN / USL_FEATS) { set_error("field grids must have 16 levels x 2 features"); return 1; }
    if (p->n > 0 && !p->x && (!p->rays_o || !p->rays_d || !p->z || p->S <= 0)) { set_error("points: need x or (rays_o, rays_d, z, S)"); return 1; }
    cudaStream_t s = (cudaStream_t)stream;
    FieldTcArgs A;
    A.f = *f; A.p = *p; A.raw = raw; A.feat = feat; A.jac = jac;
    dim3 grid((unsigned)((p->n + TC_THREADS - 1) / TC_THREADS), 2);
    if (feat) field_fwd_tc_kernel<true><<<grid, TC_CTA_THREADS, 0, s>>>(A);
    else field_fwd_tc_kernel<false><<<grid, TC_CTA_THREADS, 0, s>>>(A);
    return check_launch("usl_field_fwd (tcgen05)");
}
