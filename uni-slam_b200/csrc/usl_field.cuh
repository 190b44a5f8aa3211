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
                                            const float xB[3], uint32_t side, f32x2_t v[2][4], float w[2][3]) {
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
        for (int q = 0; q < 4; ++q) v[pnt][q] = ldg_f32x2(tab + idx[q]);      // (feature 0 | feature 1) stays a packed pair
    }
}

// part 2: bilinear in (y, z) on this lane's x side for both points, exchange of the side values inside the lane pair,
// lerp along x for the lane's own point (A on the even lane, B on the odd lane).
template <bool WITH_JAC>
__device__ __forceinline__ void pair_finish(float scale, const f32x2_t v[2][4], const float w[2][3], uint32_t side, float2 &f,
                                            float2 df[3]) {
    // every lerp acts on both features at once (FFMA2 / FADD2; each half rounds like the scalar op)
    f32x2_t b[2], by[2], bz[2];
#pragma unroll
    for (int pnt = 0; pnt < 2; ++pnt) {
        const f32x2_t w1 = pack2(w[pnt][1], w[pnt][1]), w2 = pack2(w[pnt][2], w[pnt][2]);
        const f32x2_t dy0 = fsub2(v[pnt][1], v[pnt][0]), dy1 = fsub2(v[pnt][3], v[pnt][2]);
        f32x2_t a0 = v[pnt][0], a1 = v[pnt][2];
        ffma2(a0, w1, dy0); ffma2(a1, w1, dy1);
        bz[pnt] = fsub2(a1, a0);
        b[pnt] = a0; ffma2(b[pnt], w2, bz[pnt]);
        if (WITH_JAC) { by[pnt] = dy0; ffma2(by[pnt], w2, fsub2(dy1, dy0)); }
    }
    auto xchg = [&](const f32x2_t (&t)[2], f32x2_t &s0, f32x2_t &s1) {
        const f32x2_t send = side ? t[0] : t[1];           // partner's point, my side
        const f32x2_t mine = side ? t[1] : t[0];           // my point, my side
        const f32x2_t recv = __shfl_xor_sync(0xffffffffu, send, 1);
        s0 = side ? recv : mine; s1 = side ? mine : recv;  // side-0 / side-1 value of my point
    };
    const float w0s = side ? w[1][0] : w[0][0];
    const f32x2_t w0 = pack2(w0s, w0s);
    f32x2_t s0, s1;
    xchg(b, s0, s1);
    const f32x2_t dx = fsub2(s1, s0);
    f32x2_t fp = s0; ffma2(fp, w0, dx);
    f = unpack2(fp);
    if (WITH_JAC) {
        const f32x2_t sc = pack2(scale, scale);
        f32x2_t y0, y1, z0, z1;
        xchg(by, y0, y1);
        xchg(bz, z0, z1);
        f32x2_t ty = y0, tz = z0;
        ffma2(ty, w0, fsub2(y1, y0)); ffma2(tz, w0, fsub2(z1, z0));
        df[0] = unpack2(fmul2(sc, dx)); df[1] = unpack2(fmul2(sc, ty)); df[2] = unpack2(fmul2(sc, tz));
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
#ifndef USL_FWD_PIPE
#define USL_FWD_PIPE 1
#endif
    if (LANEPAIR && UNR == 1 && USL_FWD_PIPE) {
        // software pipeline over the levels: level l's corner values are reduced to its feature (+ tangents), then the
        // gathers of level l + 1 go out, and only then the first-layer FMAs of level l issue -- a warp covers its own L2
        // latency with its own arithmetic instead of relying on the other (register-limited, 16 per SM) warps
        f32x2_t v[2][4];
        float w[2][3];
        pair_gather(g.levels[0], table, xA, xB, side, v, w);
#pragma unroll 1
        for (int l = 0; l < g.n_levels; ++l) {
            float2 f, df[3];
            pair_finish<WITH_JAC>(g.levels[l].scale, v, w, side, f, df);
            if (l + 1 < g.n_levels) pair_gather(g.levels[l + 1], table, xA, xB, side, v, w);
            accumulate(l, f, df);
        }
    } else if (LANEPAIR && UNR == 1) {
#pragma unroll 1
        for (int l = 0; l < g.n_levels; ++l) {
            f32x2_t v[2][4];
            float2 f, df[3];
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
            f32x2_t v[UNR][2][4];
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
