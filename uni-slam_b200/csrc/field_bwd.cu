// Fused field backward: decoder backward + decoder weight gradients + hash-table gradient scatter, both grids.
// Replaces loss.backward() through two tcnn.Network / nn.Linear stacks and two tcnn kernel_grid_backward launches
// (src/Mapper.py:444; src/networks/decoders.py:91-205).
//
// Shape of the kernel (sm_100a):
//   * persistent CTAs (resident CTAs per SM x SM count, half of them per grid), each looping over tiles of 128 points;
//   * the activation stash of tile i+1 (features, hidden pre-activations, clamped coordinates) and d_raw / raw arrive by
//     cp.async.bulk (TMA unit, SASS UBLKCP) into a shared-memory stage, completion on an mbarrier, while tile i issues
//     its fire-and-forget atomics: a tile is consumed into registers by the decoder backward at the START of its
//     iteration, so ONE stage suffices -- it is refilled as soon as the compute phase ends and has the whole (longer)
//     scatter phase to land.  The global-memory latency of the stash, which stalled the previous one-shot kernel for a
//     third of its time, is off the critical path and off the LSU pipe, and the small footprint (44 KB) keeps 5 CTAs
//     per SM resident;
//   * scatter by lane quads: lanes 4k..4k+3 serve the four points 4k..4k+3 together; lane bit 0 = x side, bit 1 = y side,
//     each lane owns the two z corners of its (x,y) side for all four points.  The x-neighbour corners therefore sit in
//     adjacent lanes of one RED instruction (adjacent entries of one 32-byte sector: merged by the L2), and consecutive
//     samples of a ray that fall into the SAME cell (dense coarse levels: 30-70 % of neighbouring samples) are summed in
//     registers and flushed as one atomic -- run-length aggregation without extra shuffles or shared-memory atomics;
//   * small coarse levels go to replicated private copies (L2 same-sector atomics serialise), folded afterwards.
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "usl_async.cuh"
#include "usl_field.cuh"

namespace usl {

#ifndef B2_THREADS
#define B2_THREADS 128           // = points per tile (thread = point)
#endif
#define B2_WARPS (B2_THREADS / 32)
#define B2_TILE B2_THREADS
#define B2_FROW (B2_TILE + 2)    // float2 per feature row: +16 bytes so that neighbouring levels start 4 banks apart
#define B2_HROW (B2_TILE + 4)    // floats per hidden row: likewise
#ifndef B2_MINB
#define B2_MINB 4                // resident CTAs per SM the register allocation aims for (one hidden layer)
#endif
#ifndef B2_MINB2
#define B2_MINB2 3               // ... with two hidden layers (more live registers, 26 KB of exchange tiles)
#endif

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
    unsigned int *counter;  // device work-queue counter (zero at launch; lives in the caller-zeroed scratch), or NULL: static stride
    int dbg;              // development builds (-DUSL_DEV) only: ablation switches, see usl_field_bwd
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
    B2Stage st;
    MlpSmem mlp;
    float2 w1p[USL_IN / 2][USL_HID];   // first decoder layer as (feature 0, feature 1) weight pairs per level: [level][unit]
    alignas(8) uint64_t full;
    int64_t next[2];                   // next work item of this CTA (written by thread 0, read after the CTA barrier; slot = iteration parity)
    alignas(16) float tile[1];         // [B2_WARPS][32][TROW] follows: per-point rows of decoder-gradient operands; TROW = 22 / 54
                                       // floats (one / two hidden layers): 64-bit accesses of 16 consecutive rows hit 32 distinct banks
};
template <int NH> struct B2Row { static constexpr int value = (NH == 2) ? 54 : 22; };
template <int NH> constexpr size_t b2_smem_bytes() { return offsetof(B2Smem, tile) + sizeof(float) * B2_WARPS * 32 * B2Row<NH>::value; }

__device__ __forceinline__ uint32_t stage_bytes(int cnt) { return (uint32_t)cnt * (16u * 8u + 16u * 4u + 16u + 16u + 12u); }

// Issue the loads of tile `t` of grid `gi` into the stage (one thread).
__device__ __forceinline__ void issue_tile(const FieldBwd2Args &A, B2Smem &S, int gi, int64_t t, uint64_t pol) {
    const int64_t n = A.n, i0 = t * B2_TILE;
    const int cnt = (int)min((int64_t)B2_TILE, n - i0);
    B2Stage &st = S.st;
    uint64_t *bar = &S.full;
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
__device__ __forceinline__ void load_tile_sync(const FieldBwd2Args &A, B2Smem &S, int gi, int64_t t) {
    const int64_t n = A.n, i = t * B2_TILE + threadIdx.x;
    B2Stage &st = S.st;
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

// Per-level constants hoisted out of the point loop.  Branch-free over dense / hashed levels: both index forms are two
// multiplies and combines (add or xor) followed by a wrap (conditional subtract or mask), so the kind only selects.
struct LevelConst {
    float scale;
    uint32_t size, mask, my, mz;   // multipliers of g1 / g2: (res, res^2) dense, (PRIME_Y, PRIME_Z) hashed
    bool hashed;
};
__device__ __forceinline__ LevelConst level_const(const usl_level_t &lv) {
    LevelConst c;
    c.scale = lv.scale; c.size = lv.size; c.mask = lv.size - 1u; c.hashed = lv.hashed != 0;
    c.my = c.hashed ? USL_PRIME_Y : lv.res;
    c.mz = c.hashed ? USL_PRIME_Z : lv.res * lv.res;
    return c;
}
// The two z corners on (x side sx, y side sy) of the cell of point (x0,x1,x2): entry indices, weights, packed cell id.
// Index and weight arithmetic is corner_indices<true> / corner_weights restricted to one (x,y) side: bit-identical values.
struct SideCorners2 {
    uint32_t cell;
    uint32_t i0, i1;          // entries of corner z = g2 and z = g2 + 1
    float w0, w1;
};
__device__ __forceinline__ SideCorners2 side_corners2(const LevelConst &lc, float x0, float x1, float x2, uint32_t sx, uint32_t sy) {
    SideCorners2 c;
    uint32_t g0, g1, g2;
    float f0, f1, f2;
    pos_fract(lc.scale, x0, g0, f0);
    pos_fract(lc.scale, x1, g1, f1);
    pos_fract(lc.scale, x2, g2, f2);
    c.cell = g0 | (g1 << 10) | (g2 << 20);
    const uint32_t a = g0 + sx, b = (g1 + sy) * lc.my, z0 = g2 * lc.mz, z1 = z0 + lc.mz;
    const uint32_t ab_x = a ^ b, ab_s = a + b;
    uint32_t d0 = ab_s + z0, d1 = ab_s + z1;
    d0 = (d0 >= lc.size) ? d0 - lc.size : d0;
    d1 = (d1 >= lc.size) ? d1 - lc.size : d1;
    c.i0 = lc.hashed ? ((ab_x ^ z0) & lc.mask) : d0;
    c.i1 = lc.hashed ? ((ab_x ^ z1) & lc.mask) : d1;
    float w = sx ? f0 : 1.0f - f0;                      // tcnn's multiplication order: ((f0) * f1) * f2
    w *= sy ? f1 : 1.0f - f1;
    c.w0 = w * (1.0f - f2);
    c.w1 = w * f2;
    return c;
}

// Sum the per-lane decoder-gradient accumulators over the CTA's warps and add them to the gradient tensors: one atomic
// per element per call.  Runs when a CTA leaves a grid (at most twice per CTA); the warp tiles serve as scratch.
template <int NH>
__device__ __forceinline__ void flush_decoder_grads(const usl_mlp_t &gm, const usl_mlp_t &m, float *scr, float acc[29]) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
#pragma unroll
    for (int c0 = 0; c0 < 29; c0 += 8) {
        __syncthreads();
#pragma unroll
        for (int q = 0; q < 8; ++q)
            if (c0 + q < 29) scr[(warp * 32 + lane) * 9 + q] = acc[c0 + q];
        __syncthreads();
        for (int e = tid; e < 32 * 8; e += B2_THREADS) {
            const int ln = e >> 3, q = c0 + (e & 7);
            if (q >= 29) continue;
            float sacc = 0.f;
#pragma unroll
            for (int w = 0; w < B2_WARPS; ++w) sacc += scr[(w * 32 + ln) * 9 + (e & 7)];
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
    __syncthreads();
}

template <int NH>
__global__ void __launch_bounds__(B2_THREADS, (NH == 2) ? B2_MINB2 : B2_MINB) field_bwd2_kernel(const __grid_constant__ FieldBwd2Args A) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    B2Smem &S = *reinterpret_cast<B2Smem *>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t n = A.n;
    const int64_t n_tiles = (n + B2_TILE - 1) / B2_TILE;
    const int64_t n_items = n_tiles * A.n_grids;            // work items, grid-major: item w = (grid w / n_tiles, tile w % n_tiles)
    const MlpSmem &sm = S.mlp;
    const uint64_t pol = l2_policy_evict_first();
    if (tid == 0) { mbar_init(&S.full, 1); mbar_fence_init(); }
    __syncthreads();

    // decoder weight gradients kept in registers across tiles: a1[q] = dW1 patch (packed pairs), rest[0:8] dW2 patch,
    // rest[8:10] dWo patch, rest[10:13] bias sums
    f32x2_t a1[8];
    float rest[13];
#pragma unroll
    for (int q = 0; q < 8; ++q) a1[q] = 0ull;
#pragma unroll
    for (int q = 0; q < 13; ++q) rest[q] = 0.f;

    constexpr int TROW = B2Row<NH>::value;
    float(*tile)[TROW] = reinterpret_cast<float(*)[TROW]>(S.tile + (size_t)warp * 32 * TROW);
    const uint32_t wid = (uint32_t)blockIdx.x * B2_WARPS + warp;

    auto flush = [&](int gi_) {
        float acc[29];
#pragma unroll
        for (int q = 0; q < 8; ++q) { const float2 v = unpack2(a1[q]); acc[2 * q] = v.x; acc[2 * q + 1] = v.y; a1[q] = 0ull; }
#pragma unroll
        for (int q = 0; q < 13; ++q) { acc[16 + q] = rest[q]; rest[q] = 0.f; }
        flush_decoder_grads<NH>(A.gm[gi_], A.f.mlp[gi_], S.tile, acc);
    };

    int64_t w = blockIdx.x;                                 // first item: static; later ones from the device-side queue
    if (A.bulk_ok && tid == 0 && w < n_items) issue_tile(A, S, A.gi_base + (int)(w / n_tiles), w % n_tiles, pol);
    int cur = -1;                                           // grid whose decoder is staged / whose gradients the accumulators hold
    int k = 0;
#pragma unroll 1
    for (; w < n_items; ++k) {
        const int gi = A.gi_base + (int)(w / n_tiles);
        const int64_t t = w % n_tiles;
        if (gi != cur) {                                    // at most one switch per CTA (items are grid-major)
            if (cur >= 0 && A.has_gm) flush(cur);
            __syncthreads();
            stage_mlp(A.f.mlp[gi], S.mlp);
            for (int e = tid; e < (USL_IN / 2) * USL_HID; e += B2_THREADS) {        // W1 as (feature 0, feature 1) pairs per level
                const int l = e / USL_HID, j = e % USL_HID;
                S.w1p[l][j] = make_float2(A.f.mlp[gi].w1[j * USL_IN + 2 * l], A.f.mlp[gi].w1[j * USL_IN + 2 * l + 1]);
            }
            cur = gi;
            __syncthreads();
        }
        // next item: one atomic on the work counter (dynamic load balance: items differ in cost and CTAs share SMs)
        if (tid == 0) S.next[k & 1] = A.counter ? (int64_t)atomicAdd(A.counter, 1u) + gridDim.x : w + gridDim.x;
        const usl_mlp_t &m = A.f.mlp[gi];
        const usl_grid_t &g = A.f.grid[gi];
        const int L = g.n_levels;
        if (A.bulk_ok) mbar_wait(&S.full, (uint32_t)k & 1u);
        else {
            load_tile_sync(A, S, gi, t);
            __syncthreads();
        }
        const B2Stage &st = S.st;
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
#ifdef USL_DEV
        if (A.has_gm && !(A.dbg & 4)) {
#else
        if (A.has_gm) {
#endif
            float2 *row = reinterpret_cast<float2 *>(tile[lane]);
#pragma unroll
            for (int q = 0; q < 8; ++q) row[q] = make_float2(dh1[2 * q], dh1[2 * q + 1]);
            row[8] = make_float2(du[0], du[1]); row[9] = make_float2(du[2], du[3]);
            if (NH == 2) {
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    row[10 + q] = make_float2(dh2[2 * q], dh2[2 * q + 1]);
                    row[18 + q] = make_float2(a2[2 * q], a2[2 * q + 1]);
                }
            }
            __syncwarp();
            const int pw = warp * 32;                      // first point of this warp inside the tile
            const uint32_t amask = __ballot_sync(0xffffffffu, active);   // inactive points have no stash: never touch their rows
            {   // dW1[j][k] and db1: lane = (j, parity of the level); a1[q] <-> inputs k = 2*(2q + (lane&1)) + {0,1}
                const int j = lane >> 1, lp = lane & 1;
#pragma unroll 2
                for (int pp = 0; pp < 32; ++pp) {
                    if (!((amask >> pp) & 1u)) continue;
                    const float d = tile[pp][j];
                    const f32x2_t dd = pack2(d, d);
#pragma unroll
                    for (int q = 0; q < 8; ++q)
                        ffma2(a1[q], dd, *reinterpret_cast<const f32x2_t *>(&st.F[2 * q + lp][pw + pp]));
                    if (lane < 16) rest[10] += tile[pp][lane];
                }
            }
            {   // dW2 (NH == 2): lane = (r2, parity of j), rest[q] <-> dW2[r2][2q + (lane&1)];
                // dWo: lane = (i = lane & 15, output pair lane >> 4), rest[8+e] <-> dWo[2*(lane>>4)+e][i];  db2, dbo
                const int r2 = lane >> 1, lp = lane & 1;
                const int ii = lane & 15, o0 = (lane >> 4) * 2;
#pragma unroll 2
                for (int pp = 0; pp < 32; ++pp) {
                    if (!((amask >> pp) & 1u)) continue;
                    float al;
                    if (NH == 2) {
                        const float d = tile[pp][20 + r2];
#pragma unroll
                        for (int q = 0; q < 8; ++q) rest[q] = fmaf(d, fmaxf(st.H1[2 * q + lp][pw + pp], 0.f), rest[q]);
                        if (lane < 16) rest[11] += tile[pp][20 + lane];
                        al = tile[pp][36 + ii];
                    } else {
                        al = fmaxf(st.H1[ii][pw + pp], 0.f);
                    }
                    const float2 duo = *reinterpret_cast<const float2 *>(&tile[pp][16 + o0]);
                    rest[8] = fmaf(duo.x, al, rest[8]);
                    rest[9] = fmaf(duo.y, al, rest[9]);
                    if (lane < 4) rest[12] += tile[pp][16 + lane];
                }
            }
        }

        __syncthreads();          // every warp has consumed the stage into registers: refill it while this tile scatters
        const int64_t w_next = S.next[k & 1];
        if (A.bulk_ok && tid == 0 && w_next < n_items) issue_tile(A, S, A.gi_base + (int)(w_next / n_tiles), w_next % n_tiles, pol);

        // ---- hash-table gradient scatter by lane QUADS: lanes 4k..4k+3 serve the four points 4k..4k+3 together; lane bit 0
        //      = x side, bit 1 = y side, each lane owns the two z corners of its (x,y) side for all four points.  The two x
        //      neighbours of a corner therefore sit in adjacent lanes of one RED instruction (adjacent entries, one 32-byte
        //      sector in 75 % of the cases: merged by the memory system into one atomic sector operation), and consecutive
        //      samples of a ray that fall into the SAME cell (coarse levels: 30-70 % of neighbouring samples) are summed in
        //      registers and flushed as one atomic where the run ends: 39.0 M -> 31.6 M atomic sector operations.
        //      The loop is bound by the rate at which the LSU drains the REDs (~30 cycles per warp instruction), so the index
        //      arithmetic and the level-gradient contraction d f_l = W1[:, 2l:2l+2]^T dh1 stay INSIDE it, where they are hidden;
        //      variants that moved work out of the loop (level gradients precomputed into shared memory; lane pairs with half
        //      the index arithmetic for the fine levels) measured slower (231-242 us vs 220 us).
        float2 *gt = reinterpret_cast<float2 *>(A.grad_table[gi]);
#ifdef USL_DEV
        if (gt != nullptr && !(A.dbg & 2)) {
#else
        if (gt != nullptr) {
#endif
            const int qbase = lane & ~3;
            const uint32_t sx = lane & 1u, sy = (lane >> 1) & 1u;
            float qx[4][3];
            bool qact[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
#pragma unroll
                for (int d = 0; d < 3; ++d) qx[q][d] = __shfl_sync(0xffffffffu, xc[d], qbase + q);
                qact[q] = __shfl_sync(0xffffffffu, active ? 1 : 0, qbase + q) != 0;
            }
            f32x2_t dhp[USL_HID];
#pragma unroll
            for (int j = 0; j < USL_HID; ++j) dhp[j] = pack2(dh1[j], dh1[j]);
            const int rot = (int)((wid * 5u + (uint32_t)k * 3u) % (unsigned)L);    // de-correlate the levels in flight across warps
#pragma unroll 1
            for (int it = 0; it < L; ++it) {
                int l = it + rot;
                if (l >= L) l -= L;
                f32x2_t dfp = 0ull;                                                   // (d f_l.x, d f_l.y) of this lane's own point
                const f32x2_t *wp = reinterpret_cast<const f32x2_t *>(S.w1p[l]);
#pragma unroll
                for (int j = 0; j < USL_HID; ++j) ffma2(dfp, wp[j], dhp[j]);
                const float2 df = unpack2(dfp);
                const usl_level_t &lv = g.levels[l];
                const LevelConst lc = level_const(lv);
                const uint32_t R = A.scratch ? A.rep_count[gi][l] : 1u;
                float2 *tab = (R > 1u) ? reinterpret_cast<float2 *>(A.scratch) + A.rep_offset[gi][l] + (size_t)(wid & (R - 1u)) * lv.size
                                       : gt + lv.offset;
                SideCorners2 c[4];
                float2 u0[4], u1[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float fx = __shfl_sync(0xffffffffu, df.x, qbase + q), fy = __shfl_sync(0xffffffffu, df.y, qbase + q);
                    c[q] = side_corners2(lc, qx[q][0], qx[q][1], qx[q][2], sx, sy);
                    u0[q] = make_float2(c[q].w0 * fx, c[q].w0 * fy);
                    u1[q] = make_float2(c[q].w1 * fx, c[q].w1 * fy);
                }
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    if (q > 0) {
                        const bool join = qact[q] && qact[q - 1] && (c[q].cell == c[q - 1].cell);     // continues the run of q-1
                        u0[q].x += join ? u0[q - 1].x : 0.f; u0[q].y += join ? u0[q - 1].y : 0.f;
                        u1[q].x += join ? u1[q - 1].x : 0.f; u1[q].y += join ? u1[q - 1].y : 0.f;
                    }
                    const bool ends = qact[q] && (q == 3 || !(qact[q + 1 < 4 ? q + 1 : 3] && (c[q + 1 < 4 ? q + 1 : 3].cell == c[q].cell)));
#ifdef USL_DEV
                    if (A.dbg & 1) continue;
#endif
                    if (ends) {
                        atomicAdd(tab + c[q].i0, u0[q]);
                        atomicAdd(tab + c[q].i1, u1[q]);
                    }
                }
            }
        }
        w = w_next;
    }
    if (cur >= 0 && A.has_gm) flush(cur);
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
    *n_floats = entries * 2 + 32;        // + work-queue counters (one per grid_mask value), zero-filled with the rest
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
    const bool leave_room = (grid_mask & USL_BWD_LEAVE_ROOM) != 0;
    grid_mask &= 3;
    if (grid_mask < 1 || grid_mask > 3) { set_error("usl_field_bwd: grid_mask must be 1 (sdf), 2 (colour) or 3 (both) [| USL_BWD_LEAVE_ROOM]"); return 1; }
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
    A.dbg = 0;
#ifdef USL_DEV
    { const char *e = getenv("USL_DBG_BWD"); if (e) A.dbg = atoi(e); }   // ablations: 1 no atomics, 2 no scatter phase, 4 no weight-gradient phase
    if (A.dbg & 8) A.scratch = nullptr;                                  // ablation: no replicated coarse levels
#endif
    A.counter = scratch ? reinterpret_cast<unsigned int *>(scratch + rep_entries * 2) + grid_mask : nullptr;
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
    const int64_t n_items = ((p->n + B2_TILE - 1) / B2_TILE) * A.n_grids;
    if (leave_room && per_sm > 1) --per_sm;
    int64_t nb = (int64_t)n_sm * per_sm;                      // persistent: every resident slot of the device, once
    if (nb > n_items) nb = n_items;
    const unsigned nblk = (unsigned)nb;
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
