// decode_point: one decoder (hash-grid gather + MLP + activation) evaluated on one point.
#pragma once
#include "usl_device.cuh"

namespace usl {

// Everything after the first layer's pre-activations h (and their tangents th): ReLU, optional second hidden
// layer, output layer, output activation.  out[o] activated outputs; tout[o][d] = d out / d xc.
template <bool WITH_JAC>
__device__ __forceinline__ void mlp_tail(const usl_mlp_t &m, const MlpSmem &sm, float h[USL_HID],
                                         float th[WITH_JAC ? 3 : 1][USL_HID], float out[4], float tout[4][3]) {
#pragma unroll
    for (int j = 0; j < USL_HID; ++j) {
        const bool on = h[j] > 0.f;
        h[j] = on ? h[j] : 0.f;
        if (WITH_JAC) {
#pragma unroll
            for (int d = 0; d < 3; ++d) th[d][j] = on ? th[d][j] : 0.f;
        }
    }
    float u[4], tu[4][3];
#pragma unroll
    for (int o = 0; o < 4; ++o) {
        u[o] = sm.bo[o];
        tu[o][0] = tu[o][1] = tu[o][2] = 0.f;
    }
    if (m.n_hidden == 2) {
        for (int i = 0; i < USL_HID; ++i) {
            float s = sm.b2[i], ts[3] = {0.f, 0.f, 0.f};
#pragma unroll
            for (int j = 0; j < USL_HID; ++j) {
                const float w = sm.w2[i][j];
                s = fmaf(w, h[j], s);
                if (WITH_JAC) {
#pragma unroll
                    for (int d = 0; d < 3; ++d) ts[d] = fmaf(w, th[d][j], ts[d]);
                }
            }
            if (s > 0.f) {
#pragma unroll
                for (int o = 0; o < 4; ++o) {
                    const float w = sm.wo[o][i];
                    u[o] = fmaf(w, s, u[o]);
                    if (WITH_JAC) {
#pragma unroll
                        for (int d = 0; d < 3; ++d) tu[o][d] = fmaf(w, ts[d], tu[o][d]);
                    }
                }
            }
        }
    } else {
#pragma unroll
        for (int o = 0; o < 4; ++o) {
#pragma unroll
            for (int j = 0; j < USL_HID; ++j) {
                const float w = sm.wo[o][j];
                u[o] = fmaf(w, h[j], u[o]);
                if (WITH_JAC) {
#pragma unroll
                    for (int d = 0; d < 3; ++d) tu[o][d] = fmaf(w, th[d][j], tu[o][d]);
                }
            }
        }
    }
#pragma unroll
    for (int o = 0; o < 4; ++o) {
        out[o] = act_fwd(m.out_act, u[o]);
        if (WITH_JAC) {
            const float da = act_bwd(m.out_act, out[o]);
#pragma unroll
            for (int d = 0; d < 3; ++d) tout[o][d] = da * tu[o][d];
        }
    }
}

// One decoder on one point. out[o] activated outputs; tout[o][d] = d out / d xc.
template <bool WITH_JAC, bool SAVE_FEAT, int UNR = 1, bool PAIRED = false>
__device__ __forceinline__ void decode_point(const usl_grid_t &g, const float2 *__restrict__ table,
                                             const usl_mlp_t &m, const MlpSmem &sm, const float xc[3],
                                             float2 *__restrict__ feat_out, int64_t feat_stride,
                                             float out[4], float tout[4][3], float *__restrict__ h1_out = nullptr) {
    float h[USL_HID];
    float th[WITH_JAC ? 3 : 1][USL_HID];
#pragma unroll
    for (int j = 0; j < USL_HID; ++j) {
        h[j] = sm.b1[j];
        if (WITH_JAC) { th[0][j] = 0.f; th[1][j] = 0.f; th[2][j] = 0.f; }
    }
#pragma unroll UNR
    for (int l = 0; l < g.n_levels; ++l) {
        float2 f, df[3];
        level_interp<WITH_JAC, false, PAIRED>(g.levels[l], table, xc[0], xc[1], xc[2], f, df);
        if (SAVE_FEAT && feat_out) __stcs(feat_out + (int64_t)l * feat_stride, f);    // streaming store: the stash must not evict the tables from L2
        const float4 *wa = reinterpret_cast<const float4 *>(sm.w1t[2 * l]);
        const float4 *wb = reinterpret_cast<const float4 *>(sm.w1t[2 * l + 1]);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float4 a = wa[q], b = wb[q];
            const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int j = q * 4 + e;
                h[j] = fmaf(av[e], f.x, h[j]);
                h[j] = fmaf(bv[e], f.y, h[j]);
                if (WITH_JAC) {
#pragma unroll
                    for (int d = 0; d < 3; ++d) {
                        th[d][j] = fmaf(av[e], df[d].x, th[d][j]);
                        th[d][j] = fmaf(bv[e], df[d].y, th[d][j]);
                    }
                }
            }
        }
    }
    if (SAVE_FEAT && h1_out) {   // hidden pre-activations kept for the backward pass: [16][n]
#pragma unroll
        for (int j = 0; j < USL_HID; ++j) __stcs(h1_out + (int64_t)j * feat_stride, h[j]);
    }
    mlp_tail<WITH_JAC>(m, sm, h, th, out, tout);
}

}  // namespace usl
