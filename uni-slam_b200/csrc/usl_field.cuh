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

// ---- lane-pair gather ----------------------------------------------------------------------------
// The field query is bound by the L1TEX sector-lookup rate (one 32-byte sector per clock per SM: tools/microbench.py
// measures 283 G scattered 8-byte loads/s = 0.97 per clock per SM, and field_fwd sits at 93 % of it), so what pays is fewer
// sector lookups per point.  Here lanes (2k, 2k+1) serve point A (the even lane's) and point B (the odd lane's) together:
// each lane gathers the four corners on x side (lane & 1) of BOTH points.  The two x neighbours of a corner -- adjacent
// entries of the table, the same 32-byte sector in 75 % (dense) / 50 % (hashed) of the cases -- are then fetched by adjacent
// lanes of ONE load instruction and cost one lookup instead of two.  Each lane interpolates its side bilinearly in (y, z);
// the two sides of a point meet in one exchange of the side values (2 shuffles per level, 6 with tangents) and the final
// lerp along x.  Same trilinear polynomial as level_interp (the lerp order differs: y, z, then x), same cell indices.
// part 1: cell arithmetic + the 8 gathers (4 corners of this lane's x side for A and for B); nothing waits on the loads here,
// so a caller can issue the gathers of several levels back to back before touching any value.
__device__ __forceinline__ void pair_gather(const usl_level_t &lv, const float2 *__restrict__ table, const float xA[3],
                                            const float xB[3], uint32_t side, float2 v[2][4], float w[2][3]) {
    const float2 *tab = table + lv.offset;
#pragma unroll
    for (int pnt = 0; pnt < 2; ++pnt) {
        const float *x = pnt ? xB : xA;
        const Cell c = make_cell(lv, x[0], x[1], x[2]);
        w[pnt][0] = c.w[0]; w[pnt][1] = c.w[1]; w[pnt][2] = c.w[2];
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
                if (i >= lv.size) i -= lv.size;           // clamped coordinates: one conditional subtract is the exact modulo
                idx[q] = i;
            }
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) v[pnt][q] = ldg2(tab + idx[q]);
    }
}

// part 2: bilinear in (y, z) on this lane's x side for both points, exchange of the side values inside the lane pair,
// lerp along x for the lane's own point (A on the even lane, B on the odd lane).
template <bool WITH_JAC>
__device__ __forceinline__ void pair_finish(float scale, const float2 v[2][4], const float w[2][3], uint32_t side, float2 &f,
                                            float2 df[3]) {
    float2 b[2], by[2], bz[2];
#pragma unroll
    for (int pnt = 0; pnt < 2; ++pnt) {
        const float w1 = w[pnt][1], w2 = w[pnt][2];
        const float2 dy0 = make_float2(v[pnt][1].x - v[pnt][0].x, v[pnt][1].y - v[pnt][0].y);
        const float2 dy1 = make_float2(v[pnt][3].x - v[pnt][2].x, v[pnt][3].y - v[pnt][2].y);
        const float2 a0 = make_float2(fmaf(w1, dy0.x, v[pnt][0].x), fmaf(w1, dy0.y, v[pnt][0].y));
        const float2 a1 = make_float2(fmaf(w1, dy1.x, v[pnt][2].x), fmaf(w1, dy1.y, v[pnt][2].y));
        bz[pnt] = make_float2(a1.x - a0.x, a1.y - a0.y);
        b[pnt] = make_float2(fmaf(w2, bz[pnt].x, a0.x), fmaf(w2, bz[pnt].y, a0.y));
        if (WITH_JAC) by[pnt] = make_float2(fmaf(w2, dy1.x - dy0.x, dy0.x), fmaf(w2, dy1.y - dy0.y, dy0.y));
    }
    auto xchg = [&](const float2 (&t)[2], float2 &s0, float2 &s1) {
        const float2 send = side ? t[0] : t[1];            // partner's point, my side
        const float2 mine = side ? t[1] : t[0];            // my point, my side
        const float2 recv = make_float2(__shfl_xor_sync(0xffffffffu, send.x, 1), __shfl_xor_sync(0xffffffffu, send.y, 1));
        s0 = side ? recv : mine; s1 = side ? mine : recv;  // side-0 / side-1 value of my point
    };
    const float w0 = side ? w[1][0] : w[0][0];
    float2 s0, s1;
    xchg(b, s0, s1);
    const float2 dx = make_float2(s1.x - s0.x, s1.y - s0.y);
    f = make_float2(fmaf(w0, dx.x, s0.x), fmaf(w0, dx.y, s0.y));
    if (WITH_JAC) {
        float2 y0, y1, z0, z1;
        xchg(by, y0, y1);
        xchg(bz, z0, z1);
        df[0] = make_float2(scale * dx.x, scale * dx.y);
        df[1] = make_float2(scale * fmaf(w0, y1.x - y0.x, y0.x), scale * fmaf(w0, y1.y - y0.y, y0.y));
        df[2] = make_float2(scale * fmaf(w0, z1.x - z0.x, z0.x), scale * fmaf(w0, z1.y - z0.y, z0.y));
    }
}

// One decoder on one point. out[o] activated outputs; tout[o][d] = d out / d xc.
// LANEPAIR = true (warp-collective: every lane of the warp must call it, filtered points with any in-range coordinates):
// gathers by lane pairs, see level_interp_pair.
template <bool WITH_JAC, bool SAVE_FEAT, int UNR = 1, bool PAIRED = false, bool LANEPAIR = false>
__device__ __forceinline__ void decode_point(const usl_grid_t &g, const float2 *__restrict__ table,
                                             const usl_mlp_t &m, const MlpSmem &sm, const float xc[3],
                                             float2 *__restrict__ feat_out, int64_t feat_stride,
                                             float out[4], float tout[4][3], float *__restrict__ h1_out = nullptr) {
    float xA[3], xB[3];
    const uint32_t side = threadIdx.x & 1u;
    if (LANEPAIR) {
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            const float px = __shfl_xor_sync(0xffffffffu, xc[d], 1);
            xA[d] = side ? px : xc[d];
            xB[d] = side ? xc[d] : px;
        }
    }
    // first-layer accumulators as packed pairs (units 2q, 2q+1): one FFMA2 updates two of them (same rounding as fmaf)
    f32x2_t hp[USL_HID / 2];
    f32x2_t thp[WITH_JAC ? 3 : 1][USL_HID / 2];
#pragma unroll
    for (int q = 0; q < USL_HID / 2; ++q) {
        hp[q] = pack2(sm.b1[2 * q], sm.b1[2 * q + 1]);
        if (WITH_JAC) { thp[0][q] = 0ull; thp[1][q] = 0ull; thp[2][q] = 0ull; }
    }
    auto accumulate = [&](int l, const float2 &f, const float2 df[3]) {
        if (SAVE_FEAT && feat_out) __stcs(feat_out + (int64_t)l * feat_stride, f);    // streaming store: the stash must not evict the tables from L2
        const ulonglong2 *wa = reinterpret_cast<const ulonglong2 *>(sm.w1t[2 * l]);     // weights of feature 0 / 1 of level l, units in pairs
        const ulonglong2 *wb = reinterpret_cast<const ulonglong2 *>(sm.w1t[2 * l + 1]);
        const f32x2_t fx = pack2(f.x, f.x), fy = pack2(f.y, f.y);
        f32x2_t dx[3], dy[3];
        if (WITH_JAC) {
#pragma unroll
            for (int d = 0; d < 3; ++d) { dx[d] = pack2(df[d].x, df[d].x); dy[d] = pack2(df[d].y, df[d].y); }
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const ulonglong2 a = wa[q], b = wb[q];
            ffma2(hp[2 * q], a.x, fx); ffma2(hp[2 * q], b.x, fy);
            ffma2(hp[2 * q + 1], a.y, fx); ffma2(hp[2 * q + 1], b.y, fy);
            if (WITH_JAC) {
#pragma unroll
                for (int d = 0; d < 3; ++d) {
                    ffma2(thp[d][2 * q], a.x, dx[d]); ffma2(thp[d][2 * q], b.x, dy[d]);
                    ffma2(thp[d][2 * q + 1], a.y, dx[d]); ffma2(thp[d][2 * q + 1], b.y, dy[d]);
                }
            }
        }
    };
    if (LANEPAIR && UNR == 1) {
#pragma unroll 1
        for (int l = 0; l < g.n_levels; ++l) {
            float2 v[2][4], f, df[3];
            float w[2][3];
            pair_gather(g.levels[l], table, xA, xB, side, v, w);
            pair_finish<WITH_JAC>(g.levels[l].scale, v, w, side, f, df);
            accumulate(l, f, df);
        }
    } else if (LANEPAIR) {
        // UNR levels per round: all their gathers are issued before the first value is used, so a point pays
        // n_levels / UNR memory round trips instead of n_levels.  (Measured on the mapping workload: 2 levels in flight at
        // 165 registers / 12 warps per SM 176 us, 4 levels at 245 registers / 8 warps 204 us, against 161 us for one level at
        // 122 registers / 16 warps -- resident warps hide more latency than deeper per-thread batches.)
#pragma unroll 1
        for (int l0 = 0; l0 < g.n_levels; l0 += UNR) {
            float2 v[UNR][2][4];
            float w[UNR][2][3];
#pragma unroll
            for (int u = 0; u < UNR; ++u)
                if (l0 + u < g.n_levels) pair_gather(g.levels[l0 + u], table, xA, xB, side, v[u], w[u]);
#pragma unroll
            for (int u = 0; u < UNR; ++u) {
                if (l0 + u < g.n_levels) {
                    float2 f, df[3];
                    pair_finish<WITH_JAC>(g.levels[l0 + u].scale, v[u], w[u], side, f, df);
                    accumulate(l0 + u, f, df);
                }
            }
        }
    } else {
#pragma unroll UNR
        for (int l = 0; l < g.n_levels; ++l) {
            float2 f, df[3];
            level_interp<WITH_JAC, false, PAIRED>(g.levels[l], table, xc[0], xc[1], xc[2], f, df);
            accumulate(l, f, df);
        }
    }
    float h[USL_HID];
    float th[WITH_JAC ? 3 : 1][USL_HID];
#pragma unroll
    for (int q = 0; q < USL_HID / 2; ++q) {
        const float2 v = unpack2(hp[q]);
        h[2 * q] = v.x; h[2 * q + 1] = v.y;
        if (WITH_JAC) {
#pragma unroll
            for (int d = 0; d < 3; ++d) { const float2 t = unpack2(thp[d][q]); th[d][2 * q] = t.x; th[d][2 * q + 1] = t.y; }
        }
    }
    if (SAVE_FEAT && h1_out) {   // hidden pre-activations kept for the backward pass: [16][n]
#pragma unroll
        for (int j = 0; j < USL_HID; ++j) __stcs(h1_out + (int64_t)j * feat_stride, h[j]);
    }
    mlp_tail<WITH_JAC>(m, sm, h, th, out, tout);
}


}  // namespace usl
