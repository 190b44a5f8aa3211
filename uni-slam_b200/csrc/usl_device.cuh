// Device-side building blocks shared by all kernels: hash-grid level math, decoder MLP staging.
// sm_100a only.  Reference semantics: SURVEY.md section 8a-1 (tcnn GridEncoding restated).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/unislam_b200.h"

#define USL_PRIME_Y 2654435761u
#define USL_PRIME_Z 805459861u

namespace usl {

// ---- error plumbing (host) ------------------------------------------------------------------
void set_error(const char *fmt, ...);
int check_launch(const char *what);

// ---- hash-grid level ------------------------------------------------------------------------
struct Cell {
    uint32_t g[3];
    float w[3];
};

// pos = fmaf(scale, x, 0.5f); g = floorf(pos); w = pos - g     (tcnn pos_fract, Linear interpolation)
__device__ __forceinline__ void pos_fract(float scale, float x, uint32_t &g, float &w) {
    const float pos = __fmaf_rn(scale, x, 0.5f);
    const float fl = floorf(pos);
    g = (uint32_t)(int32_t)fl;
    w = pos - fl;
}

__device__ __forceinline__ Cell make_cell(const usl_level_t &lv, float x0, float x1, float x2) {
    Cell c;
    pos_fract(lv.scale, x0, c.g[0], c.w[0]);
    pos_fract(lv.scale, x1, c.g[1], c.w[1]);
    pos_fract(lv.scale, x2, c.g[2], c.w[2]);
    return c;
}

// Entry indices (within the level) of the 8 corners; corner c: bit d set -> g_d + 1.
// tcnn grid_index<3, CoherentPrime> with `% size`: hashed levels have power-of-two size (mask).  Dense levels only
// wrap at the x == 1 boundary: for coordinates clamped to [0,1] (CLAMPED = true: every fused field kernel) the linear
// index is < 2*size, so one conditional subtract is the exact modulo; un-clamped callers (stand-alone Encoding seam)
// keep the general `%`.
template <bool CLAMPED = false>
__device__ __forceinline__ void corner_indices(const usl_level_t &lv, const Cell &c, uint32_t idx[8]) {
    if (lv.hashed) {
        const uint32_t mask = lv.size - 1u;
        const uint32_t hx0 = c.g[0], hx1 = c.g[0] + 1u;
        const uint32_t hy0 = c.g[1] * USL_PRIME_Y, hy1 = hy0 + USL_PRIME_Y;
        const uint32_t hz0 = c.g[2] * USL_PRIME_Z, hz1 = hz0 + USL_PRIME_Z;
        const uint32_t h00 = hy0 ^ hz0, h10 = hy1 ^ hz0, h01 = hy0 ^ hz1, h11 = hy1 ^ hz1;
        idx[0] = (hx0 ^ h00) & mask; idx[1] = (hx1 ^ h00) & mask;
        idx[2] = (hx0 ^ h10) & mask; idx[3] = (hx1 ^ h10) & mask;
        idx[4] = (hx0 ^ h01) & mask; idx[5] = (hx1 ^ h01) & mask;
        idx[6] = (hx0 ^ h11) & mask; idx[7] = (hx1 ^ h11) & mask;
    } else {
        const uint32_t res = lv.res, res2 = lv.res * lv.res;
        const uint32_t base = c.g[0] + c.g[1] * res + c.g[2] * res2;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            uint32_t i = base + (k & 1) + ((k >> 1) & 1) * res + ((k >> 2) & 1) * res2;
            if (CLAMPED) { if (i >= lv.size) i -= lv.size; }
            else { if (i >= lv.size) i %= lv.size; }
            idx[k] = i;
        }
    }
}

// Trilinear weights in tcnn's multiplication order: ((1 * f0) * f1) * f2.
__device__ __forceinline__ void corner_weights(const Cell &c, float wt[8]) {
    const float a0 = 1.0f - c.w[0], a1 = 1.0f - c.w[1], a2 = 1.0f - c.w[2];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        float w = (k & 1) ? c.w[0] : a0;
        w *= (k & 2) ? c.w[1] : a1;
        w *= (k & 4) ? c.w[2] : a2;
        wt[k] = w;
    }
}

__device__ __forceinline__ float2 ldg2(const float2 *p) { return __ldg(p); }

// Gradient scatter of one level: 8 corners = 4 x-pairs.  The two corners of an x-pair land in the same
// 16-byte slot whenever their indices differ only in bit 0 (dense levels: even base index; CoherentPrime hash:
// even g_x, because the x coefficient is 1) -- then ONE 16-byte vector atomic replaces two 8-byte ones.
// L2 atomic throughput on B200 is per 32-byte sector operation (tools/microbench.py), so this removes 25 % of them.
__device__ __forceinline__ void scatter_level(float2 *__restrict__ tab, const uint32_t idx[8], const float wt[8],
                                              float dfx, float dfy) {
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        const uint32_t i0 = idx[2 * p], i1 = idx[2 * p + 1];
        const float2 v0 = make_float2(wt[2 * p] * dfx, wt[2 * p] * dfy);
        const float2 v1 = make_float2(wt[2 * p + 1] * dfx, wt[2 * p + 1] * dfy);
        if ((i0 ^ i1) == 1u) {
            const bool lo_first = (i0 & 1u) == 0u;
            const float4 v = lo_first ? make_float4(v0.x, v0.y, v1.x, v1.y) : make_float4(v1.x, v1.y, v0.x, v0.y);
            atomicAdd(reinterpret_cast<float4 *>(tab + (i0 & ~1u)), v);
        } else {
            atomicAdd(tab + i0, v0);
            atomicAdd(tab + i1, v1);
        }
    }
}

// Half-cell scatter for the lane-pair scheme of field_bwd: this lane serves the 4 corners on x-side `side` (0: g_x,
// 1: g_x + 1) of the point at (x0,x1,x2).  Lanes (2k, 2k+1) call it with the SAME point and side = lane & 1, so corner q of
// both sides -- neighbouring entries, one 32-byte sector in 75 % of the cases -- travels in one instruction and the L2
// merges the pair into one atomic sector operation.  Only the point's coordinates and its two level gradients cross
// the lane pair (2 shuffles per level); indices and weights are computed where they are used.  Index and weight
// arithmetic is corner_indices<true> / corner_weights restricted to one x-side (bit-identical values).
__device__ __forceinline__ void scatter_level_side(float2 *__restrict__ tab, const usl_level_t &lv, float x0, float x1, float x2,
                                                   uint32_t side, float dfx, float dfy, bool act) {
    const Cell c = make_cell(lv, x0, x1, x2);
    uint32_t idx[4];
    if (lv.hashed) {
        const uint32_t mask = lv.size - 1u;
        const uint32_t hx = c.g[0] + side;
        const uint32_t hy0 = c.g[1] * USL_PRIME_Y, hy1 = hy0 + USL_PRIME_Y;
        const uint32_t hz0 = c.g[2] * USL_PRIME_Z, hz1 = hz0 + USL_PRIME_Z;
        idx[0] = (hx ^ hy0 ^ hz0) & mask; idx[1] = (hx ^ hy1 ^ hz0) & mask;
        idx[2] = (hx ^ hy0 ^ hz1) & mask; idx[3] = (hx ^ hy1 ^ hz1) & mask;
    } else {
        const uint32_t res = lv.res, res2 = lv.res * lv.res;
        const uint32_t base = c.g[0] + side + c.g[1] * res + c.g[2] * res2;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            uint32_t i = base + (q & 1) * res + (q >> 1) * res2;
            if (i >= lv.size) i -= lv.size;               // clamped coordinates: one conditional subtract is the exact modulo
            idx[q] = i;
        }
    }
    const float wx = side ? c.w[0] : 1.0f - c.w[0];
    const float a1 = 1.0f - c.w[1], a2 = 1.0f - c.w[2];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        float w = wx;                                      // tcnn's multiplication order: ((f0) * f1) * f2
        w *= (q & 1) ? c.w[1] : a1;
        w *= (q & 2) ? c.w[2] : a2;
        if (act) atomicAdd(tab + idx[q], make_float2(w * dfx, w * dfy));
    }
}

// Gather the 8 corners of one level and interpolate; optionally the d f / d x tangents.
// TCNN_ORDER = true reproduces tcnn's corner loop (8 weights, fma chain) for the stand-alone Encoding seam;
// false evaluates the same trilinear polynomial as nested lerps (x, then y, then z): 14 ops per feature instead
// of ~26, and the tangents fall out of the differences already formed (6+2+0 extra ops). Results agree to a few
// ulp (the <=1e-6 feature tolerance of the parity gate); cell indices are identical by construction.
// PAIRED = true (warp-collective, every lane of the warp must call it): lanes (2k, 2k+1) fetch ONE point's x-pair per
// load instruction, so the two corners -- same 32-byte sector in 75 % of the cases -- are served by one L1 wavefront /
// L2 sector request; the values are handed back to their owner with two shuffles.
template <bool WITH_JAC, bool TCNN_ORDER = false, bool PAIRED = false>
__device__ __forceinline__ void level_interp(const usl_level_t &lv, const float2 *__restrict__ table,
                                             float x0, float x1, float x2, float2 &f, float2 df[3]) {
    const Cell c = make_cell(lv, x0, x1, x2);
    uint32_t idx[8];
    corner_indices<!TCNN_ORDER>(lv, c, idx);      // the fused kernels (TCNN_ORDER = false) always pass clamped coordinates
    const float2 *tab = table + lv.offset;
    float2 v[8];
    // (128-bit loads for x-pairs that share a 16-byte slot were measured: divergent 200 us, predicated 183 us vs
    // 182 us for plain loads -- the gather is bound by the L1 data pipe + latency, not by sector lookups -- so the
    // gather stays 8 x 64-bit; the same pairing does pay off for the atomics, see scatter_level.)
    if (PAIRED) {
        const bool odd = (threadIdx.x & 1) != 0;
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            const uint32_t i0 = idx[2 * p], i1 = idx[2 * p + 1];
            const uint32_t pi1 = __shfl_xor_sync(0xffffffffu, i1, 1);
            const float2 a = ldg2(tab + (odd ? pi1 : i0));     // the even lane's point: (own corner 2p | partner's corner 2p+1)
            const float2 b = ldg2(tab + (odd ? i0 : pi1));     // the odd lane's point
            const float2 give = odd ? a : b;                   // the partner's second corner, fetched on its behalf
            v[2 * p] = odd ? b : a;
            v[2 * p + 1] = make_float2(__shfl_xor_sync(0xffffffffu, give.x, 1), __shfl_xor_sync(0xffffffffu, give.y, 1));
        }
    } else {
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = ldg2(tab + idx[k]);
    }
    if (TCNN_ORDER) {
        float wt[8];
        corner_weights(c, wt);
        f.x = 0.f; f.y = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            f.x = __fmaf_rn(wt[k], v[k].x, f.x);
            f.y = __fmaf_rn(wt[k], v[k].y, f.y);
        }
    }
    if (!TCNN_ORDER || WITH_JAC) {
        const float w0 = c.w[0], w1 = c.w[1], w2 = c.w[2];
        float2 dx[4], a[4];
#pragma unroll
        for (int p = 0; p < 4; ++p) {                       // along x; p = y + 2 z
            dx[p] = make_float2(v[2 * p + 1].x - v[2 * p].x, v[2 * p + 1].y - v[2 * p].y);
            a[p] = make_float2(fmaf(w0, dx[p].x, v[2 * p].x), fmaf(w0, dx[p].y, v[2 * p].y));
        }
        float2 dy[2], b[2], bx[2];
#pragma unroll
        for (int z = 0; z < 2; ++z) {                       // along y
            dy[z] = make_float2(a[2 * z + 1].x - a[2 * z].x, a[2 * z + 1].y - a[2 * z].y);
            b[z] = make_float2(fmaf(w1, dy[z].x, a[2 * z].x), fmaf(w1, dy[z].y, a[2 * z].y));
            if (WITH_JAC)
                bx[z] = make_float2(fmaf(w1, dx[2 * z + 1].x - dx[2 * z].x, dx[2 * z].x), fmaf(w1, dx[2 * z + 1].y - dx[2 * z].y, dx[2 * z].y));
        }
        const float2 dz = make_float2(b[1].x - b[0].x, b[1].y - b[0].y);
        if (!TCNN_ORDER) f = make_float2(fmaf(w2, dz.x, b[0].x), fmaf(w2, dz.y, b[0].y));
        if (WITH_JAC) {
            const float s = lv.scale;
            df[0] = make_float2(s * fmaf(w2, bx[1].x - bx[0].x, bx[0].x), s * fmaf(w2, bx[1].y - bx[0].y, bx[0].y));
            df[1] = make_float2(s * fmaf(w2, dy[1].x - dy[0].x, dy[0].x), s * fmaf(w2, dy[1].y - dy[0].y, dy[0].y));
            df[2] = make_float2(s * dz.x, s * dz.y);
        }
    }
}

// ---- decoder weights staged in shared memory --------------------------------------------------
// Layout (floats): w1t[32][16] (transposed: [k][j]), b1[16], w2[16][16], b2[16], wo[4][16], bo[4]
struct alignas(16) MlpSmem {
    float w1t[USL_IN][USL_HID];
    float b1[USL_HID];
    float w2[USL_HID][USL_HID];
    float b2[USL_HID];
    float wo[4][USL_HID];
    float bo[4];
};

__device__ __forceinline__ void stage_mlp(const usl_mlp_t &m, MlpSmem &s) {
    const int t = threadIdx.x, nt = blockDim.x;
    for (int i = t; i < USL_IN * USL_HID; i += nt) {
        const int j = i / USL_IN, k = i % USL_IN;          // w1 is [j][k] row-major
        s.w1t[k][j] = m.w1[i];
    }
    for (int i = t; i < USL_HID; i += nt) {
        s.b1[i] = m.b1 ? m.b1[i] : 0.f;
        s.b2[i] = (m.n_hidden == 2 && m.b2) ? m.b2[i] : 0.f;
    }
    for (int i = t; i < USL_HID * USL_HID; i += nt)
        (&s.w2[0][0])[i] = (m.n_hidden == 2) ? m.w2[i] : 0.f;
    for (int i = t; i < 4 * USL_HID; i += nt)
        (&s.wo[0][0])[i] = (i / USL_HID < m.n_out) ? m.wo[i] : 0.f;
    for (int i = t; i < 4; i += nt) s.bo[i] = (i < m.n_out && m.bo) ? m.bo[i] : 0.f;
}

__device__ __forceinline__ float act_fwd(int act, float u) {
    if (act == USL_ACT_TANH) return tanhf(u);
    if (act == USL_ACT_SIGMOID) return 1.0f / (1.0f + expf(-u));
    return u;
}
// derivative expressed with the activated output y
__device__ __forceinline__ float act_bwd(int act, float y) {
    if (act == USL_ACT_TANH) return 1.0f - y * y;
    if (act == USL_ACT_SIGMOID) return y * (1.0f - y);
    return 1.0f;
}

// Normalised coordinate of sample s on ray r: x = ((o + d*z) - lo) / (hi - lo), op-for-op like
// src/utils/Renderer.py:132,137 (separate mul/add, IEEE division; no FMA contraction).
__device__ __forceinline__ float norm_coord(float o, float d, float z, float lo, float hi) {
    const float p = __fadd_rn(o, __fmul_rn(d, z));
    return __fdiv_rn(__fsub_rn(p, lo), __fsub_rn(hi, lo));
}

// pose [qw,qx,qy,qz,tx,ty,tz] -> rows 0..2 of c2w (3x4 row-major): pytorch3d quaternion_to_matrix for an
// un-normalised real-first quaternion (src/common.py:196-208), op-for-op like torch eager (no FMA contraction).
__device__ __forceinline__ void pose_to_c2w(const float *__restrict__ pose, float o[12]) {
    const float r = pose[0], i = pose[1], j = pose[2], kk = pose[3];
#define USL_MUL(a, b) __fmul_rn(a, b)
#define USL_ADD(a, b) __fadd_rn(a, b)
#define USL_SUB(a, b) __fsub_rn(a, b)
    const float two_s = __fdiv_rn(2.0f, USL_ADD(USL_ADD(USL_ADD(USL_MUL(r, r), USL_MUL(i, i)), USL_MUL(j, j)), USL_MUL(kk, kk)));
    o[0] = USL_SUB(1.0f, USL_MUL(two_s, USL_ADD(USL_MUL(j, j), USL_MUL(kk, kk))));
    o[1] = USL_MUL(two_s, USL_SUB(USL_MUL(i, j), USL_MUL(kk, r)));
    o[2] = USL_MUL(two_s, USL_ADD(USL_MUL(i, kk), USL_MUL(j, r)));
    o[3] = pose[4];
    o[4] = USL_MUL(two_s, USL_ADD(USL_MUL(i, j), USL_MUL(kk, r)));
    o[5] = USL_SUB(1.0f, USL_MUL(two_s, USL_ADD(USL_MUL(i, i), USL_MUL(kk, kk))));
    o[6] = USL_MUL(two_s, USL_SUB(USL_MUL(j, kk), USL_MUL(i, r)));
    o[7] = pose[5];
    o[8] = USL_MUL(two_s, USL_SUB(USL_MUL(i, kk), USL_MUL(j, r)));
    o[9] = USL_MUL(two_s, USL_ADD(USL_MUL(j, kk), USL_MUL(i, r)));
    o[10] = USL_SUB(1.0f, USL_MUL(two_s, USL_ADD(USL_MUL(i, i), USL_MUL(j, j))));
    o[11] = pose[6];
#undef USL_MUL
#undef USL_ADD
#undef USL_SUB
}

// ---- packed FP32 (Blackwell FFMA2: two IEEE fp32 FMAs per instruction, each rounded like fmaf) ----
typedef unsigned long long f32x2_t;
__device__ __forceinline__ f32x2_t pack2(float lo, float hi) {
    f32x2_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ float2 unpack2(f32x2_t v) {
    float2 r;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
    return r;
}
__device__ __forceinline__ void ffma2(f32x2_t &acc, f32x2_t a, f32x2_t b) {       // acc = a * b + acc (element-wise)
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(a), "l"(b));
}

__device__ __forceinline__ f32x2_t fsub2(f32x2_t a, f32x2_t b) {                    // a - b (element-wise, round to nearest)
    f32x2_t r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2_t fmul2(f32x2_t a, f32x2_t b) {
    f32x2_t r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2_t lerp2(f32x2_t w, f32x2_t a, f32x2_t b) {          // fmaf(w, b - a, a) on both halves
    f32x2_t r = a;
    ffma2(r, w, fsub2(b, a));
    return r;
}
__device__ __forceinline__ f32x2_t ldg_f32x2(const void *p) {                        // 8-byte read-only load straight into a packed pair
    return __ldg(reinterpret_cast<const unsigned long long *>(p));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace usl
