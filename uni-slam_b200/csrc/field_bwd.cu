// Fused field backward: decoder backward + decoder weight gradients + hash-table gradient scatter, both grids.
// Replaces loss.backward() through two tcnn.Network / nn.Linear stacks and two tcnn kernel_grid_backward launches
// (src/Mapper.py:444; src/networks/decoders.py:91-205).
//
// Shape of the kernel (sm_100a):
//   * persistent CTAs (resident CTAs per SM x SM count, half of them per grid), each looping over tiles of 128 points;
//   * the activation stash of tile i+1 (features, hidden pre-activations, clamped coordinates) and d_raw / raw arrive by
//     cp.async.bulk (TMA unit, SASS UBLKCP) into a double-buffered shared-memory stage, completion on an mbarrier, while
//     tile i runs its decoder backward and issues its fire-and-forget atomics: the global-memory latency of the stash,
//     which stalled the previous one-shot kernel for a third of its time, is off the critical path and off the LSU pipe;
//   * scatter by lane quads: lanes 4k..4k+3 serve the four points 4k..4k+3 together; lane bit 0 = x side, bit 1 = y side,
//     each lane owns the two z corners of its (x,y) side for all four points.  The x-neighbour corners therefore sit in
//     adjacent lanes of one RED instruction (adjacent entries of one 32-byte sector: merged by the L2), and consecutive
//     samples of a ray that fall into the SAME cell (dense coarse levels: 30-70 % of neighbouring samples) are summed in
//     registers and flushed as one atomic -- run-length aggregation without extra shuffles or shared-memory atomics;
//   * small coarse levels go to replicated private copies (L2 same-sector atomics serialise), folded afterwards.
#include <cstddef>
#include <cstdio>
#include <cstring>

#include "usl_async.cuh"
#include "usl_field.cuh"

namespace usl {

#define B2_THREADS 128
#define B2_WARPS (B2_THREADS / 32)
#define B2_TILE 128
#define B2_FROW (B2_TILE + 2)    // float2 per feature row: +16 bytes so that neighbouring levels start 4 banks apart
#define B2_HROW (B2_TILE + 4)    // floats per hidden row: likewise

struct FieldBwd2Args {
    usl_field_t f;
    int64_t n;
    const float *raw;     // [n,4] saved outputs
    const float *feat;    // stash: [2][L][n][2] features, [2][16][n] hidden pre-activations, [3][n] clamped coordinates
    const float *d_raw;   // [n,4]
    float *grad_table[2];
    usl_mlp_t gm[2];
    int has_gm;
    int gi_base, n_grids;
    float *scratch;
    uint32_t rep_count[2][USL_MAX_LEVELS];
    uint32_t rep_offset[2][USL_MAX_LEVELS];
    int bulk_ok;          // 1: every tile row is 16-byte aligned and sized -> cp.async.bulk; 0: plain cooperative loads
};

// One pipeline stage = everything the backward needs about one tile of 128 points of one grid.
struct alignas(128) B2Stage {
    float2 F[USL_IN / 2][B2_FROW];     // interpolated features, [level][point]          16 KB
    float H1[USL_HID][B2_HROW];        // hidden pre-activations, [unit][point]           8 KB
    float4 DR[B2_TILE];                // d_raw rows                                      2 KB
    float4 RW[B2_TILE];                // raw rows                                        2 KB
    float X[3][B2_TILE];               // clamped normalised coordinates (x0 < 0: point inactive)  1.5 KB
};

struct B2Smem {
    B2Stage st[2];
    MlpSmem mlp;
    alignas(8) uint64_t full[2];
    alignas(16) float tile[1];         // [B2_WARPS][32][TROW] follows (TROW = 20 floats for one hidden layer, 52 for two)
};
template <int NH> struct B2Row { static constexpr int value = (NH == 2) ? 52 : 20; };
template <int NH> constexpr size_t b2_smem_bytes() { return offsetof(B2Smem, tile) + sizeof(float) * B2_WARPS * 32 * B2Row<NH>::value; }

__device__ __forceinline__ uint32_t stage_bytes(int cnt) { return (uint32_t)cnt * (16u * 8u + 16u * 4u + 16u + 16u + 12u); }

// Issue the loads of tile `t` of grid `gi` into stage `s` (one thread).
__device__ __forceinline__ void issue_tile(const FieldBwd2Args &A, B2Smem &S, int s, int gi, int64_t t, uint64_t pol) {
    const int64_t n = A.n, i0 = t * B2_TILE;
    const int cnt = (int)min((int64_t)B2_TILE, n - i0);
    B2Stage &st = S.st[s];
    uint64_t *bar = &S.full[s];
    mbar_arrive_expect_tx(bar, stage_bytes(cnt));
    const float2 *feat = reinterpret_cast<const float2 *>(A.feat) + ((int64_t)gi * (USL_IN / 2)) * n + i0;
#pragma unroll 1
    for (int l = 0; l < USL_IN / 2; ++l) bulk_g2s(st.F[l], feat + (int64_t)l * n, (uint32_t)cnt * 8u, bar, pol);
    const float *h1 = A.feat + (int64_t)2 * USL_IN * n + ((int64_t)gi * USL_HID) * n + i0;
#pragma unroll 1
    for (int j = 0; j < USL_HID; ++j) bulk_g2s(st.H1[j], h1 + (int64_t)j * n, (uint32_t)cnt * 4u, bar, pol);
    const float *xs = A.feat + (int64_t)2 * (USL_IN + USL_HID) * n + i0;
#pragma unroll 1
    for (int d = 0; d < 3; ++d) bulk_g2s(st.X[d], xs + (int64_t)d * n, (uint32_t)cnt * 4u, bar, pol);
    bulk_g2s(st.DR, A.d_raw + i0 * 4, (uint32_t)cnt * 16u, bar, pol);
    bulk_g2s(st.RW, A.raw + i0 * 4, (uint32_t)cnt * 16u, bar, pol);
}

// Fallback when the rows are not 16-byte aligned / sized (n % 4 != 0 or odd base pointers): every thread fetches its own
// column with ordinary loads.  Same stage layout, so everything downstream is shared.
__device__ __forceinline__ void load_tile_sync(const FieldBwd2Args &A, B2Smem &S, int s, int gi, int64_t t) {
    const int64_t n = A.n, i = t * B2_TILE + threadIdx.x;
    B2Stage &st = S.st[s];
    const int p = threadIdx.x;
    if (i < n) {
        const float2 *feat = reinterpret_cast<const float2 *>(A.feat) + ((int64_t)gi * (USL_IN / 2)) * n + i;
#pragma unroll
        for (int l = 0; l < USL_IN / 2; ++l) st.F[l][p] = __ldcs(feat + (int64_t)l * n);
        const float *h1 = A.feat + (int64_t)2 * USL_IN * n + ((int64_t)gi * USL_HID) * n + i;
#pragma unroll
        for (int j = 0; j < USL_HID; ++j) st.H1[j][p] = __ldcs(h1 + (int64_t)j * n);
        const float *xs = A.feat + (int64_t)2 * (USL_IN + USL_HID) * n + i;
#pragma unroll
        for (int d = 0; d < 3; ++d) st.X[d][p] = __ldcs(xs + (int64_t)d * n);
        st.DR[p] = make_float4(A.d_raw[i * 4], A.d_raw[i * 4 + 1], A.d_raw[i * 4 + 2], A.d_raw[i * 4 + 3]);
        st.RW[p] = make_float4(A.raw[i * 4], A.raw[i * 4 + 1], A.raw[i * 4 + 2], A.raw[i * 4 + 3]);
    }
}

// The two z corners on (x side sx, y side sy) of the cell of point (x0,x1,x2) at level lv: entry indices and weights.
// Index and weight arithmetic is corner_indices<true> / corner_weights restricted to one (x,y) side: bit-identical.
struct SideCorners {
    uint32_t g0, g1, g2;      // cell (for the same-cell test)
    uint32_t i0, i1;          // entries of corner z = g2 and z = g2 + 1
    float w0, w1;
};
__device__ __forceinline__ SideCorners side_corners(const usl_level_t &lv, float x0, float x1, float x2, uint32_t sx, uint32_t sy) {
    SideCorners c;
    float f0, f1, f2;
    pos_fract(lv.scale, x0, c.g0, f0);
    pos_fract(lv.scale, x1, c.g1, f1);
    pos_fract(lv.scale, x2, c.g2, f2);
    if (lv.hashed) {
        const uint32_t mask = lv.size - 1u;
        const uint32_t h = (c.g0 + sx) ^ ((c.g1 + sy) * USL_PRIME_Y);
        const uint32_t hz = c.g2 * USL_PRIME_Z;
        c.i0 = (h ^ hz) & mask;
        c.i1 = (h ^ (hz + USL_PRIME_Z)) & mask;
    } else {
        const uint32_t res = lv.res, res2 = lv.res * lv.res;
        uint32_t b = (c.g0 + sx) + (c.g1 + sy) * res + c.g2 * res2;
        uint32_t b1 = b + res2;
        if (b >= lv.size) b -= lv.size;               // clamped coordinates: one conditional subtract is the exact modulo
        if (b1 >= lv.size) b1 -= lv.size;
        c.i0 = b; c.i1 = b1;
    }
    float w = sx ? f0 : 1.0f - f0;                    // tcnn's multiplication order: ((f0) * f1) * f2
    w *= sy ? f1 : 1.0f - f1;
    c.w0 = w * (1.0f - f2);
    c.w1 = w * f2;
    return c;
}

template <int NH>
__global__ void __launch_bounds__(B2_THREADS, 3) field_bwd2_kernel(const __grid_constant__ FieldBwd2Args A) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    B2Smem &S = *reinterpret_cast<B2Smem *>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int gi = A.gi_base + (A.n_grids == 2 ? (int)(blockIdx.x & 1u) : 0);
    const int cta = (A.n_grids == 2) ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
    const int n_cta = (A.n_grids == 2) ? (int)(gridDim.x >> 1) : (int)gridDim.x;
    const usl_mlp_t &m = A.f.mlp[gi];
    const usl_grid_t &g = A.f.grid[gi];
    const int L = g.n_levels;
    const int64_t n = A.n;
    const int64_t n_tiles = (n + B2_TILE - 1) / B2_TILE;

    stage_mlp(m, S.mlp);
    if (tid == 0) { mbar_init(&S.full[0], 1); mbar_init(&S.full[1], 1); mbar_fence_init(); }
    __syncthreads();
    const MlpSmem &sm = S.mlp;
    const uint64_t pol = l2_policy_evict_first();
    if (A.bulk_ok && tid == 0 && cta < n_tiles) issue_tile(A, S, 0, gi, cta, pol);

    float acc1[16], acc2[8], acco[2], accb[3] = {0.f, 0.f, 0.f};      // decoder weight gradients, kept across tiles
#pragma unroll
    for (int q = 0; q < 16; ++q) acc1[q] = 0.f;
#pragma unroll
    for (int q = 0; q < 8; ++q) acc2[q] = 0.f;
    acco[0] = acco[1] = 0.f;

    constexpr int TROW = B2Row<NH>::value;
    float(*tile)[TROW] = reinterpret_cast<float(*)[TROW]>(S.tile + (size_t)warp * 32 * TROW);
    float2 *gt = reinterpret_cast<float2 *>(A.grad_table[gi]);
    const uint32_t wid = (uint32_t)cta * B2_WARPS + warp;
    const uint32_t sx = lane & 1u, sy = (lane >> 1) & 1u;
    const int qbase = lane & ~3;

    int k = 0;
#pragma unroll 1
    for (int64_t t = cta; t < n_tiles; t += n_cta, ++k) {
        const int s = k & 1;
        if (A.bulk_ok) {
            // stage s^1 was last read in iteration k-1, before its __syncthreads: free to refill
            if (tid == 0 && t + n_cta < n_tiles) issue_tile(A, S, s ^ 1, gi, t + n_cta, pol);
            mbar_wait(&S.full[s], (uint32_t)(k >> 1) & 1u);
        } else {
            load_tile_sync(A, S, s, gi, t);
            __syncthreads();
        }
        const B2Stage &st = S.st[s];
        const int cnt = (int)min((int64_t)B2_TILE, n - t * B2_TILE);
        const int p = tid;
        float xc[3];
        xc[0] = st.X[0][p]; xc[1] = st.X[1][p]; xc[2] = st.X[2][p];
        const bool active = (p < cnt) && (xc[0] >= 0.f);
        if (!active) { xc[0] = xc[1] = xc[2] = 0.f; }

        // ---- decoder backward on the point's own column of the stage ----
        float h1[USL_HID];
#pragma unroll
        for (int j = 0; j < USL_HID; ++j) h1[j] = active ? st.H1[j][p] : 0.f;
        float du[4] = {0.f, 0.f, 0.f, 0.f};
        if (active) {
            const float4 dr = st.DR[p], rw = st.RW[p];
            if (gi == 0) du[0] = dr.w * act_bwd(m.out_act, rw.w);
            else {
                du[0] = dr.x * act_bwd(m.out_act, rw.x);
                du[1] = dr.y * act_bwd(m.out_act, rw.y);
                du[2] = dr.z * act_bwd(m.out_act, rw.z);
            }
        }
        float dh1[USL_HID];
        float a2[USL_HID], dh2[USL_HID];
        if (NH == 2) {
#pragma unroll
            for (int j = 0; j < USL_HID; ++j) dh1[j] = 0.f;
#pragma unroll
            for (int q = 0; q < USL_HID; ++q) {
                float sacc = sm.b2[q];
#pragma unroll
                for (int j = 0; j < USL_HID; ++j) sacc = fmaf(sm.w2[q][j], fmaxf(h1[j], 0.f), sacc);
                float d = 0.f;
#pragma unroll
                for (int o = 0; o < 4; ++o) d = fmaf(sm.wo[o][q], du[o], d);
                d = (sacc > 0.f) ? d : 0.f;
                a2[q] = fmaxf(sacc, 0.f);
                dh2[q] = d;
#pragma unroll
                for (int j = 0; j < USL_HID; ++j) dh1[j] = fmaf(sm.w2[q][j], d, dh1[j]);
            }
#pragma unroll
            for (int j = 0; j < USL_HID; ++j) dh1[j] = (h1[j] > 0.f) ? dh1[j] : 0.f;
        } else {
#pragma unroll
            for (int j = 0; j < USL_HID; ++j) {
                float d = 0.f;
#pragma unroll
                for (int o = 0; o < 4; ++o) d = fmaf(sm.wo[o][j], du[o], d);
                dh1[j] = (h1[j] > 0.f) ? d : 0.f;
            }
        }

        // ---- decoder weight gradients: each lane owns a patch of every matrix, summed over the warp's 32 points ----
        // tile row (per point): [0:16] dh1, [16:20] du, [20:36] dh2, [36:52] a2 (the last two only when NH == 2);
        // features and first-layer activations are read straight from the stage.
        if (A.has_gm) {
            float4 *row = reinterpret_cast<float4 *>(tile[lane]);
#pragma unroll
            for (int q = 0; q < 4; ++q) row[q] = make_float4(dh1[4 * q], dh1[4 * q + 1], dh1[4 * q + 2], dh1[4 * q + 3]);
            row[4] = make_float4(du[0], du[1], du[2], du[3]);
            if (NH == 2) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    row[5 + q] = make_float4(dh2[4 * q], dh2[4 * q + 1], dh2[4 * q + 2], dh2[4 * q + 3]);
                    row[9 + q] = make_float4(a2[4 * q], a2[4 * q + 1], a2[4 * q + 2], a2[4 * q + 3]);
                }
            }
            __syncwarp();
            const int pw = warp * 32;                      // first point of this warp inside the tile
            const uint32_t amask = __ballot_sync(0xffffffffu, active);   // inactive points have no stash: never touch their rows
            {   // dW1[j][k] and db1: lane = (j, parity of the level); acc1[2q+f] <-> input k = 2*(2q + (lane&1)) + f
                const int j = lane >> 1, lp = lane & 1;
#pragma unroll 2
                for (int pp = 0; pp < 32; ++pp) {
                    if (!((amask >> pp) & 1u)) continue;
                    const float d = tile[pp][j];
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const float2 v = st.F[2 * q + lp][pw + pp];
                        acc1[2 * q] = fmaf(d, v.x, acc1[2 * q]);
                        acc1[2 * q + 1] = fmaf(d, v.y, acc1[2 * q + 1]);
                    }
                    if (lane < 16) accb[0] += tile[pp][lane];
                }
            }
            {   // dW2 (NH == 2): lane = (r2, parity of j), acc2[q] <-> dW2[r2][2q + (lane&1)];
                // dWo: lane = (i = lane & 15, output pair lane >> 4), acco[e] <-> dWo[2*(lane>>4)+e][i];  db2, dbo
                const int r2 = lane >> 1, lp = lane & 1;
                const int ii = lane & 15, o0 = (lane >> 4) * 2;
#pragma unroll 2
                for (int pp = 0; pp < 32; ++pp) {
                    if (!((amask >> pp) & 1u)) continue;
                    float al;
                    if (NH == 2) {
                        const float d = tile[pp][20 + r2];
#pragma unroll
                        for (int q = 0; q < 8; ++q) acc2[q] = fmaf(d, fmaxf(st.H1[2 * q + lp][pw + pp], 0.f), acc2[q]);
                        if (lane < 16) accb[1] += tile[pp][20 + lane];
                        al = tile[pp][36 + ii];
                    } else {
                        al = fmaxf(st.H1[ii][pw + pp], 0.f);
                    }
                    const float2 duo = *reinterpret_cast<const float2 *>(&tile[pp][16 + o0]);
                    acco[0] = fmaf(duo.x, al, acco[0]);
                    acco[1] = fmaf(duo.y, al, acco[1]);
                    if (lane < 4) accb[2] += tile[pp][16 + lane];
                }
            }
        }
        __syncthreads();          // every warp is done with stage s: the producer may refill it two iterations from now

        // ---- hash-table gradient scatter by lane quads ----
        if (gt != nullptr) {
            float qx[4][3];
            bool qact[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
#pragma unroll
                for (int d = 0; d < 3; ++d) qx[q][d] = __shfl_sync(0xffffffffu, xc[d], qbase + q);
                qact[q] = __shfl_sync(0xffffffffu, active ? 1 : 0, qbase + q) != 0;
            }
            const int rot = (int)((wid * 5u + (uint32_t)k * 3u) % (unsigned)L);    // de-correlate the levels in flight across warps
#pragma unroll 1
            for (int it = 0; it < L; ++it) {
                int l = it + rot;
                if (l >= L) l -= L;
                float dfx = 0.f, dfy = 0.f;
                const float4 *wa = reinterpret_cast<const float4 *>(sm.w1t[2 * l]);
                const float4 *wb = reinterpret_cast<const float4 *>(sm.w1t[2 * l + 1]);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float4 a = wa[q], b = wb[q];
                    dfx = fmaf(a.x, dh1[4 * q], dfx); dfx = fmaf(a.y, dh1[4 * q + 1], dfx);
                    dfx = fmaf(a.z, dh1[4 * q + 2], dfx); dfx = fmaf(a.w, dh1[4 * q + 3], dfx);
                    dfy = fmaf(b.x, dh1[4 * q], dfy); dfy = fmaf(b.y, dh1[4 * q + 1], dfy);
                    dfy = fmaf(b.z, dh1[4 * q + 2], dfy); dfy = fmaf(b.w, dh1[4 * q + 3], dfy);
                }
                const usl_level_t &lv = g.levels[l];
                const uint32_t R = A.scratch ? A.rep_count[gi][l] : 1u;
                float2 *tab = (R > 1u) ? reinterpret_cast<float2 *>(A.scratch) + A.rep_offset[gi][l] + (size_t)(wid & (R - 1u)) * lv.size
                                       : gt + lv.offset;
                // walk the quad's four points; a run of points in the same cell is summed and flushed once
                uint32_t pi0 = 0, pi1 = 0, pg0 = 0, pg1 = 0, pg2 = 0;
                float2 v0 = make_float2(0.f, 0.f), v1 = v0;
                bool pending = false;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float fx = __shfl_sync(0xffffffffu, dfx, qbase + q), fy = __shfl_sync(0xffffffffu, dfy, qbase + q);
                    const SideCorners c = side_corners(lv, qx[q][0], qx[q][1], qx[q][2], sx, sy);
                    const bool same = pending && qact[q] && (c.g0 == pg0) && (c.g1 == pg1) && (c.g2 == pg2);
                    if (pending && !same) {
                        atomicAdd(tab + pi0, v0);
                        atomicAdd(tab + pi1, v1);
                        pending = false;
                    }
                    if (qact[q]) {
                        if (same) {
                            v0.x = fmaf(c.w0, fx, v0.x); v0.y = fmaf(c.w0, fy, v0.y);
                            v1.x = fmaf(c.w1, fx, v1.x); v1.y = fmaf(c.w1, fy, v1.y);
                        } else {
                            v0 = make_float2(c.w0 * fx, c.w0 * fy); v1 = make_float2(c.w1 * fx, c.w1 * fy);
                            pi0 = c.i0; pi1 = c.i1; pg0 = c.g0; pg1 = c.g1; pg2 = c.g2;
                            pending = true;
                        }
                    }
                }
                if (pending) {
                    atomicAdd(tab + pi0, v0);
                    atomicAdd(tab + pi1, v1);
                }
            }
        }
    }

    // ---- block reduction of the decoder gradients, one atomic per element per CTA (once, after the last tile) ----
    if (A.has_gm) {
        __syncthreads();
        float *red = reinterpret_cast<float *>(&S.st[0]);          // stages are idle now: [B2_WARPS][32][33] floats
        float *mine = red + (warp * 32 + lane) * 33;                // stride 33: conflict-free
#pragma unroll
        for (int q = 0; q < 16; ++q) mine[q] = acc1[q];
#pragma unroll
        for (int q = 0; q < 8; ++q) mine[16 + q] = acc2[q];
        mine[24] = acco[0]; mine[25] = acco[1];
        mine[26] = accb[0]; mine[27] = accb[1]; mine[28] = accb[2];
        __syncthreads();
        const usl_mlp_t &gm = A.gm[gi];
        for (int e = tid; e < 32 * 29; e += B2_THREADS) {
            const int ln = e / 29, q = e % 29;
            float sacc = 0.f;
#pragma unroll
            for (int w = 0; w < B2_WARPS; ++w) sacc += red[(w * 32 + ln) * 33 + q];
            if (q < 16) {                                   // dW1[j][k]: lane ln = (j, level parity), q = 2*(level >> 1) + feature
                const int j = ln >> 1, kk = 4 * (q >> 1) + 2 * (ln & 1) + (q & 1);
                if (gm.w1) atomicAdd(gm.w1 + j * USL_IN + kk, sacc);
            } else if (q < 24) {                            // dW2[r2][2*(q-16) + parity]
                if (NH == 2 && gm.w2) atomicAdd(gm.w2 + (ln >> 1) * USL_HID + 2 * (q - 16) + (ln & 1), sacc);
            } else if (q < 26) {                            // dWo[2*(ln>>4) + (q-24)][ln & 15]
                const int o = (ln >> 4) * 2 + (q - 24), ii = ln & 15;
                if (o < m.n_out && gm.wo) atomicAdd(gm.wo + o * USL_HID + ii, sacc);
            } else if (q == 26) {
                if (ln < 16 && gm.b1) atomicAdd(gm.b1 + ln, sacc);
            } else if (q == 27) {
                if (NH == 2 && ln < 16 && gm.b2) atomicAdd(gm.b2 + ln, sacc);
            } else {
                if (ln < m.n_out && gm.bo) atomicAdd(gm.bo + ln, sacc);
            }
        }
    }
}

// ---- replicated coarse levels ---------------------------------------------------------------------
// L2 atomic throughput collapses on small tables (tools/microbench_footprint.py: 79 G ops/s on 32 KB, 121 G on
// 175 KB, 220 G from 4 MB up) because operations on one 32-byte sector serialise.  The coarse dense levels are
// exactly such tables and every sample hits them, so the scatter writes them into R private copies (picked by warp
// id, ~1 MB per level in total) and a tiny second kernel folds the copies into the gradient table.
static void plan_replicas(const usl_field_t *f, uint32_t cnt[2][USL_MAX_LEVELS], uint32_t off[2][USL_MAX_LEVELS], int64_t *total_entries) {
    int64_t o = 0;
    const uint64_t max_bytes = 512u * 1024u, target = 1024u * 1024u;
    for (int gi = 0; gi < 2; ++gi)
        for (int l = 0; l < USL_MAX_LEVELS; ++l) {
            cnt[gi][l] = 1; off[gi][l] = 0;
            if (l >= f->grid[gi].n_levels) continue;
            const uint64_t bytes = (uint64_t)f->grid[gi].levels[l].size * 8u;
            if (bytes >= max_bytes) continue;
            uint32_t r = 1;
            while (r < 32u && (uint64_t)r * bytes < target) r *= 2;
            cnt[gi][l] = r; off[gi][l] = (uint32_t)o;
            o += (int64_t)r * f->grid[gi].levels[l].size;
        }
    *total_entries = o;
}

__global__ void __launch_bounds__(256) fold_replicas_kernel(const __grid_constant__ FieldBwd2Args A) {
    const int gi = A.gi_base + blockIdx.y;
    float2 *gt = reinterpret_cast<float2 *>(A.grad_table[gi]);
    if (!gt) return;
    uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    const usl_grid_t &g = A.f.grid[gi];
    for (int l = 0; l < g.n_levels; ++l) {
        const uint32_t R = A.rep_count[gi][l], sz = g.levels[l].size;
        if (R <= 1u) continue;
        if (e < sz) {
            const float2 *src = reinterpret_cast<const float2 *>(A.scratch) + A.rep_offset[gi][l] + e;
            float sx = 0.f, sy = 0.f;
            for (uint32_t r = 0; r < R; ++r) { const float2 v = src[(size_t)r * sz]; sx += v.x; sy += v.y; }
            float2 *dst = gt + g.levels[l].offset + e;
            float2 cur = *dst;
            cur.x += sx; cur.y += sy;
            *dst = cur;
            return;
        }
        e -= sz;
    }
}

}  // namespace usl

using namespace usl;

extern "C" {

int usl_field_stash_floats(int64_t n_points, int64_t *n_floats) {
    if (!n_floats || n_points < 0) { set_error("usl_field_stash_floats: bad arguments"); return 1; }
    *n_floats = n_points * (2 * (USL_IN + USL_HID) + 3);
    return 0;
}

int usl_field_bwd_scratch_floats(const usl_field_t *f, int64_t *n_floats) {
    if (!f || !n_floats) { set_error("usl_field_bwd_scratch_floats: null argument"); return 1; }
    uint32_t cnt[2][USL_MAX_LEVELS], off[2][USL_MAX_LEVELS];
    int64_t entries = 0;
    plan_replicas(f, cnt, off, &entries);
    *n_floats = entries * 2;
    return 0;
}

int usl_field_bwd(const usl_field_t *f, const usl_points_t *p, const float *raw, const float *feat,
                  const float *d_raw, float *grad_table_sdf, float *grad_table_rgb, const usl_mlp_t *gm,
                  float *scratch, int grid_mask, usl_stream_t stream) {
    if (!f || !p) { set_error("usl_field_bwd: null field/points"); return 1; }
    for (int gi = 0; gi < 2; ++gi) {
        if (f->grid[gi].n_levels != USL_IN / USL_FEATS) { set_error("field grids must have 16 levels x 2 features"); return 1; }
        if (f->mlp[gi].n_hidden < 1 || f->mlp[gi].n_hidden > 2 || f->mlp[gi].n_out < 1 || f->mlp[gi].n_out > 3) {
            set_error("unsupported decoder shape"); return 1;
        }
    }
    if (p->n <= 0) return 0;
    if (!feat || !raw || !d_raw) { set_error("usl_field_bwd: raw, feat and d_raw are required"); return 1; }
    if (f->mlp[0].n_hidden != f->mlp[1].n_hidden) { set_error("usl_field_bwd: decoders must share n_hidden"); return 1; }
    if (grid_mask < 1 || grid_mask > 3) { set_error("usl_field_bwd: grid_mask must be 1 (sdf), 2 (colour) or 3 (both)"); return 1; }
    FieldBwd2Args A;
    A.f = *f; A.n = p->n; A.raw = raw; A.feat = feat; A.d_raw = d_raw;
    A.grad_table[0] = grad_table_sdf; A.grad_table[1] = grad_table_rgb;
    A.has_gm = gm ? 1 : 0;
    if (gm) { A.gm[0] = gm[0]; A.gm[1] = gm[1]; }
    A.gi_base = (grid_mask == 2) ? 1 : 0;
    A.n_grids = (grid_mask == 3) ? 2 : 1;
    // bulk copies need 16-byte aligned, 16-byte sized rows: n % 4 == 0 makes every row of every tile so
    const uintptr_t align_or = (uintptr_t)raw | (uintptr_t)feat | (uintptr_t)d_raw;
    A.bulk_ok = ((p->n % 4) == 0 && (align_or & 15u) == 0) ? 1 : 0;
    int64_t rep_entries = 0;
    plan_replicas(f, A.rep_count, A.rep_offset, &rep_entries);
    A.scratch = (rep_entries > 0) ? scratch : nullptr;
    cudaStream_t s = (cudaStream_t)stream;

    int dev = 0, n_sm = 0, per_sm = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    const bool nh2 = f->mlp[0].n_hidden == 2;
    const size_t smem = nh2 ? b2_smem_bytes<2>() : b2_smem_bytes<1>();
    const void *fn = nh2 ? (const void *)field_bwd2_kernel<2> : (const void *)field_bwd2_kernel<1>;
    if (cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
        cudaGetLastError(); set_error("usl_field_bwd: cannot reserve %zu bytes of shared memory", smem); return 1;
    }
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, B2_THREADS, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
    const int64_t n_tiles = (p->n + B2_TILE - 1) / B2_TILE;
    int64_t ctas_per_grid = ((int64_t)n_sm * per_sm) / A.n_grids;
    if (ctas_per_grid < 1) ctas_per_grid = 1;
    if (ctas_per_grid > n_tiles) ctas_per_grid = n_tiles;
    const unsigned nblk = (unsigned)(ctas_per_grid * A.n_grids);
    if (nh2) field_bwd2_kernel<2><<<nblk, B2_THREADS, smem, s>>>(A);
    else field_bwd2_kernel<1><<<nblk, B2_THREADS, smem, s>>>(A);
    if (check_launch("usl_field_bwd")) return 1;
    if (A.scratch) {
        uint32_t per_grid = 0;
        for (int gi = A.gi_base; gi < A.gi_base + A.n_grids; ++gi) {
            uint32_t t = 0;
            for (int l = 0; l < f->grid[gi].n_levels; ++l) if (A.rep_count[gi][l] > 1) t += f->grid[gi].levels[l].size;
            if (t > per_grid) per_grid = t;
        }
        if (per_grid) fold_replicas_kernel<<<dim3((per_grid + 255) / 256, A.n_grids), 256, 0, s>>>(A);
        return check_launch("usl_field_bwd (fold)");
    }
    return 0;
}

}  // extern "C"
