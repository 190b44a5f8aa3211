/*
 * oracle/grid_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C CPU restatement of the tiny-cuda-nn multi-resolution HashGrid encoding
 * (NVlabs/tiny-cuda-nn @ 2ec562e853e6f482b5d09168705205f46358fb39, pinned by the
 * reference at requirements.txt:90; constructed at src/UNISLAM.py:224-259 with
 * otype=HashGrid, n_levels=16, n_features_per_level=2, base_resolution=16,
 * interpolation=Linear, hash=CoherentPrime, dtype=float; called at
 * src/networks/decoders.py:101-103).
 *
 * PARITY UNPINNED for the tcnn arithmetic: tiny-cuda-nn's source is not in
 * /root/reference (un-vendored dependency, CUDA-only, no network), and the
 * reference ships no golden vectors.  The algorithm below is tcnn's published
 * GridEncoding (include/tiny-cuda-nn/encodings/grid.h, common_device.h) restated
 * from SURVEY.md section 8a-1; the per-level tables are pinned against the
 * known-answer tables in BASELINE.md section 2.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this library.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

#define ORC_MAX_LEVELS 32
#define ORC_F 2 /* n_features_per_level (UNISLAM.py:226 level_dim=2) */

typedef struct {
    float scale;        /* exp2f(l*log2f(pls))*base - 1            */
    uint32_t res;       /* ceilf(scale)+1                          */
    uint32_t size;      /* entries in this level (T_l)             */
    uint32_t offset;    /* first entry of the level in the table   */
    uint32_t hashed;    /* 1 if res^3 > size (index goes via hash) */
} orc_level_t;

/* tcnn GridEncoding constructor: offsets table.  SURVEY 8a-1. */
int orc_grid_levels(int n_levels, int log2_hashmap_size, int base_resolution,
                    double per_level_scale, orc_level_t *out, uint32_t *total_entries) {
    if (n_levels <= 0 || n_levels > ORC_MAX_LEVELS) return -1;
    /* tcnn stores per_level_scale as float and takes std::log2 of it in float */
    const float pls = (float)per_level_scale;
    const float log2_pls = log2f(pls);
    uint32_t offset = 0;
    for (int l = 0; l < n_levels; ++l) {
        const float scale = exp2f((float)l * log2_pls) * (float)base_resolution - 1.0f;
        const uint32_t res = (uint32_t)ceilf(scale) + 1u;
        const uint32_t max_params = 0xFFFFFFFFu / 2u;
        const double dense = (double)res * (double)res * (double)res;
        uint32_t n = dense > (double)max_params ? max_params : (uint32_t)dense;
        n = (n + 7u) / 8u * 8u;                                   /* next_multiple(.,8) */
        const uint32_t cap = 1u << log2_hashmap_size;
        if (n > cap) n = cap;
        out[l].scale = scale;
        out[l].res = res;
        out[l].size = n;
        out[l].offset = offset;
        out[l].hashed = (dense > (double)n) ? 1u : 0u;
        offset += n;
    }
    if (total_entries) *total_entries = offset;
    return 0;
}

/* tcnn grid_index<3, CoherentPrime>.  uint32 wrap-around arithmetic throughout. */
static inline uint32_t orc_index(const orc_level_t *lv, uint32_t gx, uint32_t gy, uint32_t gz) {
    const uint32_t g[3] = {gx, gy, gz};
    uint32_t stride = 1u, index = 0u;
    for (int d = 0; d < 3 && stride <= lv->size; ++d) {
        index += g[d] * stride;
        stride *= lv->res;
    }
    if (lv->size < stride) {
        index = (gx * 1u) ^ (gy * 2654435761u) ^ (gz * 805459861u);
    }
    return index % lv->size;
}

uint32_t orc_grid_index(const orc_level_t *lv, uint32_t gx, uint32_t gy, uint32_t gz) {
    return orc_index(lv, gx, gy, gz);
}

/* pos_fract: pos = fmaf(scale, x, 0.5f); g = floorf(pos); w = pos - g */
static inline void orc_pos_fract(float scale, float x, uint32_t *g, float *w) {
    float pos = fmaf(scale, x, 0.5f);
    float fl = floorf(pos);
    *g = (uint32_t)(int32_t)fl;
    *w = pos - fl;
}

/*
 * Hash / table indices of the 8 corners, all levels.
 * x: [n,3] in [0,1]; idx_out: [n, n_levels, 8] (entry index WITHIN the level, before
 * adding the level offset); w_out (optional): [n, n_levels, 8] trilinear weights.
 * Corner c: bit d of c selects g_d+1 (weight w_d) else g_d (weight 1-w_d).
 */
void orc_grid_corners(const orc_level_t *lv, int n_levels, const float *x, int64_t n,
                      uint32_t *idx_out, float *w_out) {
    for (int64_t i = 0; i < n; ++i) {
        for (int l = 0; l < n_levels; ++l) {
            uint32_t g[3]; float w[3];
            for (int d = 0; d < 3; ++d) orc_pos_fract(lv[l].scale, x[i * 3 + d], &g[d], &w[d]);
            for (int c = 0; c < 8; ++c) {
                float wt = 1.0f; uint32_t gl[3];
                for (int d = 0; d < 3; ++d) {
                    if (c & (1 << d)) { wt *= w[d]; gl[d] = g[d] + 1u; }
                    else { wt *= 1.0f - w[d]; gl[d] = g[d]; }
                }
                idx_out[(i * n_levels + l) * 8 + c] = orc_index(&lv[l], gl[0], gl[1], gl[2]);
                if (w_out) w_out[(i * n_levels + l) * 8 + c] = wt;
            }
        }
    }
}

/* Forward: y[n, n_levels*2] = trilinear interpolation of the fp32 table. */
void orc_grid_encode_fwd(const orc_level_t *lv, int n_levels, const float *params,
                         const float *x, int64_t n, float *y) {
    for (int64_t i = 0; i < n; ++i) {
        for (int l = 0; l < n_levels; ++l) {
            uint32_t g[3]; float w[3];
            for (int d = 0; d < 3; ++d) orc_pos_fract(lv[l].scale, x[i * 3 + d], &g[d], &w[d]);
            float acc[ORC_F] = {0.f, 0.f};
            const float *tab = params + (size_t)lv[l].offset * ORC_F;
            for (int c = 0; c < 8; ++c) {
                float wt = 1.0f; uint32_t gl[3];
                for (int d = 0; d < 3; ++d) {
                    if (c & (1 << d)) { wt *= w[d]; gl[d] = g[d] + 1u; }
                    else { wt *= 1.0f - w[d]; gl[d] = g[d]; }
                }
                const uint32_t e = orc_index(&lv[l], gl[0], gl[1], gl[2]);
                for (int f = 0; f < ORC_F; ++f) acc[f] = fmaf(wt, tab[(size_t)e * ORC_F + f], acc[f]);
            }
            for (int f = 0; f < ORC_F; ++f) y[i * n_levels * ORC_F + l * ORC_F + f] = acc[f];
        }
    }
}

/* Backward wrt table: grad[e,f] += w * dy (double accumulation to separate order noise). */
void orc_grid_encode_bwd_params(const orc_level_t *lv, int n_levels, const float *x,
                                const float *dy, int64_t n, double *grad) {
    for (int64_t i = 0; i < n; ++i) {
        for (int l = 0; l < n_levels; ++l) {
            uint32_t g[3]; float w[3];
            for (int d = 0; d < 3; ++d) orc_pos_fract(lv[l].scale, x[i * 3 + d], &g[d], &w[d]);
            double *tab = grad + (size_t)lv[l].offset * ORC_F;
            for (int c = 0; c < 8; ++c) {
                float wt = 1.0f; uint32_t gl[3];
                for (int d = 0; d < 3; ++d) {
                    if (c & (1 << d)) { wt *= w[d]; gl[d] = g[d] + 1u; }
                    else { wt *= 1.0f - w[d]; gl[d] = g[d]; }
                }
                const uint32_t e = orc_index(&lv[l], gl[0], gl[1], gl[2]);
                for (int f = 0; f < ORC_F; ++f)
                    tab[(size_t)e * ORC_F + f] += (double)wt * (double)dy[i * n_levels * ORC_F + l * ORC_F + f];
            }
        }
    }
}

/* Backward wrt input: dx[d] = sum_{l,f} dy[l,f] * scale_l * sum_{other corners} w_other*(v(g_d+1)-v(g_d)) */
void orc_grid_encode_bwd_input(const orc_level_t *lv, int n_levels, const float *params,
                               const float *x, const float *dy, int64_t n, float *dx) {
    for (int64_t i = 0; i < n; ++i) {
        double acc[3] = {0, 0, 0};
        for (int l = 0; l < n_levels; ++l) {
            uint32_t g[3]; float w[3];
            for (int d = 0; d < 3; ++d) orc_pos_fract(lv[l].scale, x[i * 3 + d], &g[d], &w[d]);
            const float *tab = params + (size_t)lv[l].offset * ORC_F;
            for (int gd = 0; gd < 3; ++gd) {
                for (int c = 0; c < 4; ++c) {
                    float wt = lv[l].scale; uint32_t gl[3];
                    for (int nd = 0; nd < 2; ++nd) {
                        const int d = nd >= gd ? nd + 1 : nd;
                        if (c & (1 << nd)) { wt *= w[d]; gl[d] = g[d] + 1u; }
                        else { wt *= 1.0f - w[d]; gl[d] = g[d]; }
                    }
                    gl[gd] = g[gd];
                    const uint32_t e0 = orc_index(&lv[l], gl[0], gl[1], gl[2]);
                    gl[gd] = g[gd] + 1u;
                    const uint32_t e1 = orc_index(&lv[l], gl[0], gl[1], gl[2]);
                    for (int f = 0; f < ORC_F; ++f) {
                        const float diff = tab[(size_t)e1 * ORC_F + f] - tab[(size_t)e0 * ORC_F + f];
                        acc[gd] += (double)wt * (double)diff * (double)dy[i * n_levels * ORC_F + l * ORC_F + f];
                    }
                }
            }
        }
        for (int d = 0; d < 3; ++d) dx[i * 3 + d] = (float)acc[d];
    }
}
