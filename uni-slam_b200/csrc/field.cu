// Fused field query: clamp -> hash-grid gather (16 levels) -> decoder MLP -> activation, per point,
// and its backward (decoder weight gradients + hash-table gradient scatter).
// Replaces Decoders.forward / get_raw_sdf / get_raw_rgb (src/networks/decoders.py:91-205), i.e. two
// tcnn.Encoding calls + two tcnn.Network / nn.Linear stacks, together with the point construction
// and normalisation of src/utils/Renderer.py:132-137.  blockIdx.y selects the grid (0 sdf, 1 colour)
// so each thread carries one decoder's registers and the two grids run as independent CTAs.
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "usl_field.cuh"

namespace usl {

struct FieldArgs {
    usl_field_t f;
    usl_points_t p;
    float *raw;   // [n,4]
    float *feat;  // [2][L][n][2]
    float *jac;   // [n,12]
    float *sdf;   // [n]   (sdf-only mode)
};

// Loads point i: clamped normalised coordinate xc, clamp gate (1 inside [0,1], else 0).
// oi (optional): index of the point in the ray-major output arrays (differs from i only in sample-major mode).
__device__ __forceinline__ bool load_point(const usl_points_t &p, const usl_field_t &f, int64_t i, float xc[3],
                                           float gate[3], int64_t *oi = nullptr) {
    float x[3];
    if (oi) *oi = i;
    if (p.x) {
#pragma unroll
        for (int d = 0; d < 3; ++d) x[d] = p.x[i * 3 + d];
    } else {
        int64_t r = i / p.S, zi = i;
        if (p.sample_major) {
            const int64_t R = p.n / p.S;
            r = i % R;
            zi = r * p.S + i / R;
            if (oi) *oi = zi;
        }
        if (p.valid && !p.valid[r]) return false;
        const float z = p.z[zi];
#pragma unroll
        for (int d = 0; d < 3; ++d) x[d] = norm_coord(p.rays_o[r * 3 + d], p.rays_d[r * 3 + d], z, f.bound_lo[d], f.bound_hi[d]);
    }
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        xc[d] = fminf(fmaxf(x[d], 0.f), 1.f);                  // torch.clamp(p_nor, 0, 1), decoders.py:101
        gate[d] = (x[d] >= 0.f && x[d] <= 1.f) ? 1.f : 0.f;   // clamp backward passes grad on the closed interval
    }
    return true;
}

template <bool WITH_JAC, bool SAVE_FEAT>
#ifndef USL_FWD_THREADS
#define USL_FWD_THREADS 256
#endif
#ifndef USL_FWD_MINB
#define USL_FWD_MINB 2
#endif
#ifndef USL_FWD_UNR
#define USL_FWD_UNR 1            // levels whose gathers are in flight together (with the Jacobian)
#endif
__global__ void __launch_bounds__(USL_FWD_THREADS, (WITH_JAC ? USL_FWD_MINB : (1024 / USL_FWD_THREADS))) field_fwd_kernel(const __grid_constant__ FieldArgs A) {
    __shared__ MlpSmem sm;
    const int gi = blockIdx.y;
    stage_mlp(A.f.mlp[gi], sm);
    __syncthreads();
    float xc[3] = {0.f, 0.f, 0.f}, gate[3] = {0.f, 0.f, 0.f};
    int64_t ti = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;   // thread's point; i below = its slot in the ray-major outputs
    const bool in_range = ti < A.p.n;
    const bool active = in_range && load_point(A.p, A.f, ti, xc, gate, &ti);
    if (SAVE_FEAT && in_range) {
        // clamped coordinates for the backward pass, component-major [3][n]; x0 = -1 marks a filtered point (the only part
        // of the stash written for those), so the backward never has to re-read rays / z / valid
        float *xs = A.feat + (int64_t)2 * (USL_IN + USL_HID) * A.p.n + ti;
        if (gi == 0) {
            __stcs(xs, active ? xc[0] : -1.0f);
            __stcs(xs + A.p.n, xc[1]);
            __stcs(xs + 2 * A.p.n, xc[2]);
        }
    }
    if (!__any_sync(0xffffffffu, active)) return;           // warps made only of filtered rays cost nothing
    const int64_t i = ti;
    const usl_grid_t &g = A.f.grid[gi];
    float out[4], tout[4][3];
    // stash layout: features [2][L][n][2] then hidden pre-activations [2][16][n]
    float2 *fo = (SAVE_FEAT && active) ? reinterpret_cast<float2 *>(A.feat) + ((int64_t)gi * g.n_levels) * A.p.n + i : nullptr;
    float *ho = (SAVE_FEAT && active) ? A.feat + (int64_t)2 * USL_IN * A.p.n + ((int64_t)gi * USL_HID) * A.p.n + i : nullptr;
    // PAIRED gather (template arg 4) measured 165.6 vs 163.5 us un-paired: the forward is latency / occupancy bound, not
    // bound by sector requests -- left off.  (The same lane pairing is what speeds up the atomics in field_bwd.)
#ifndef USL_FWD_LANEPAIR
#define USL_FWD_LANEPAIR 1
#endif
    decode_point<WITH_JAC, SAVE_FEAT, USL_FWD_UNR, false, (USL_FWD_LANEPAIR != 0)>(g, reinterpret_cast<const float2 *>(A.f.table[gi]), A.f.mlp[gi],
                                                                        sm, xc, fo, A.p.n, out, tout, ho);
    if (!active) return;
    if (gi == 0) {
        A.raw[i * 4 + 3] = out[0];
        if (WITH_JAC) {
#pragma unroll
            for (int d = 0; d < 3; ++d) __stcs(A.jac + (int64_t)(9 + d) * A.p.n + i, tout[0][d] * gate[d]);   // component-major: coalesced
        }
    } else {
#pragma unroll
        for (int o = 0; o < 3; ++o) {
            A.raw[i * 4 + o] = out[o];
            if (WITH_JAC) {
#pragma unroll
                for (int d = 0; d < 3; ++d) __stcs(A.jac + (int64_t)(o * 3 + d) * A.p.n + i, tout[o][d] * gate[d]);
            }
        }
    }
}

__global__ void __launch_bounds__(256) field_sdf_kernel(const __grid_constant__ FieldArgs A) {
    __shared__ MlpSmem sm;
    stage_mlp(A.f.mlp[0], sm);
    __syncthreads();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= A.p.n) return;
    float xc[3], gate[3];
    if (!load_point(A.p, A.f, i, xc, gate)) return;
    float out[4], tout[4][3];
    decode_point<false, false>(A.f.grid[0], reinterpret_cast<const float2 *>(A.f.table[0]), A.f.mlp[0], sm, xc,
                               nullptr, 0, out, tout);
    A.sdf[i] = out[0];
}

// ---- dense SDF query for meshing (Mesher.get_grid_uniform + eval_points, Mesher.py:134-195) ----
struct QueryArgs {
    usl_field_t f;
    const float *ax, *ay, *az;
    int nx, ny, nz, y_begin, y_end;
    float *out;
};

// CTA = tile of 32 (x) x 2 * QP_WARPS (z) grid points of one y row; lane = x, and every thread owns a PAIR of consecutive
// z.  Lane = x because the dense levels store entries x-fastest and the hash's x coefficient is 1: the 32 lanes of a gather
// hit neighbouring entries (a few 32-byte sectors) instead of the 32 sectors a z-line of lanes would touch.
// The one-point-per-thread form of this kernel was issue- and latency-bound, not byte-bound (ncu: 10.3 G warp instructions
// for 132 M points, 74 % of the issue rate, L1TEX lookups at 46 % of their ceiling), so this form spends fewer instructions
// per point and hides its own latency:
//   * along z the two points of a pair sit in the same or in neighbouring cells of most levels (mesh spacing <= cell), so
//     per level the thread does the x / y cell arithmetic and the (x, y) half of the hash once and gathers z-PLANES (4
//     corners, collapsed bilinearly in x then y) instead of cells: 2 planes when the pair shares a cell, 3 when the second
//     point is in the next cell, 4 otherwise -- against 4 planes' worth for independent points; each point finishes with one
//     lerp along z; interpolation in packed pairs (feature 0 | feature 1: FFMA2 / FADD2);
//   * software pipeline over the levels: the gathers of level l + 1 are issued before the 64 first-layer FMAs of level l;
//   * the first-layer weights of a level are read once for both points.
// The lerp order (x, y, z), the cell indices and the FMA order of the first layer are those of decode_point: the values are
// bit-identical to the one-point-per-thread form.  Measured on the 132 M-point Replica volume: 12.3 ms one point per thread,
// 11.4 with 4 levels of gathers in flight, 10.6 / 9.9 ms for runs of 4 / 2 z without the pipeline (164 / 122 registers),
// 9.4 ms for this kernel at 125 registers / 4 CTAs per SM; 10.3 ms at 3 or at 5 CTAs per SM (137 registers / 96 with spills),
// 13.0 at 6.
#define QT_X 32

// z-planes of one level for one thread.  Keys and offsets are BYTE offsets into the level (entry index << 3), so that one
// LOP3 (hashed) or one add + conditional subtract (dense) yields the load offset.
template <bool HASHED>
struct PlaneKey {
    uint32_t k[4];        // hashed: ((gx + dx) ^ ((gy + dy) * PRIME_Y)) << 3; dense: (gx + dx + (gy + dy) * res) << 3
    uint32_t zstep;       // hashed: PRIME_Z << 3; dense: (res * res) << 3
    uint32_t lim;         // hashed: (size - 1) << 3 (mask); dense: size << 3
    __device__ __forceinline__ PlaneKey(const usl_level_t &lv, uint32_t gx, uint32_t gy) {
        if (HASHED) {
            const uint32_t hy0 = gy * USL_PRIME_Y, hy1 = hy0 + USL_PRIME_Y;
            k[0] = (gx ^ hy0) << 3; k[1] = ((gx + 1u) ^ hy0) << 3; k[2] = (gx ^ hy1) << 3; k[3] = ((gx + 1u) ^ hy1) << 3;
            zstep = USL_PRIME_Z << 3; lim = (lv.size - 1u) << 3;
        } else {
            const uint32_t b = gx + gy * lv.res;
            k[0] = b << 3; k[1] = (b + 1u) << 3; k[2] = (b + lv.res) << 3; k[3] = (b + lv.res + 1u) << 3;
            zstep = (lv.res * lv.res) << 3; lim = lv.size << 3;
        }
    }
    // the four corners of plane gz (zoff = gz * zstep), as packed (feature 0, feature 1) pairs
    __device__ __forceinline__ void gather(const char *__restrict__ tab, uint32_t zoff, f32x2_t v[4]) const {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            uint32_t o;
            if (HASHED) o = (k[q] ^ zoff) & lim;                       // shifting commutes with xor / and: same entry as grid_index
            else { o = k[q] + zoff; if (o >= lim) o -= lim; }          // clamped coordinates: one conditional subtract is the exact modulo
            v[q] = ldg_f32x2(tab + o);
        }
    }
};

__device__ __forceinline__ f32x2_t plane_collapse(const f32x2_t v[4], f32x2_t wx, f32x2_t wy) {   // lerp x, then y (level_interp's order)
    return lerp2(wy, lerp2(wx, v[0], v[1]), lerp2(wx, v[2], v[3]));
}

// Plane slots of a level: A = plane(gz0), B = plane(gz0 + 1) always; C = plane(gz1 + 1) unless both points share a cell;
// D = plane(gz1) only when the second point is more than one cell away.
#ifndef QP_WARPS
#define QP_WARPS 4
#endif
#ifndef QP_MINB
#define QP_MINB 4
#endif
struct PairLevel {               // what the consuming half of an iteration needs about the level whose values are in flight
    f32x2_t wx2, wy2, wz0, wz1;
    bool same, adj;
};

template <bool HASHED>
__device__ __forceinline__ void pair_issue(const usl_level_t &lv, const char *__restrict__ tab, uint32_t gx, uint32_t gy, uint32_t gz0,
                                           uint32_t gz1, bool same, bool adj, f32x2_t (*v)[4]) {
    const PlaneKey<HASHED> K(lv, gx, gy);
    const uint32_t z0 = gz0 * K.zstep;
    K.gather(tab, z0, v[0]);
    K.gather(tab, z0 + K.zstep, v[1]);
    if (!same) {
        const uint32_t z1 = gz1 * K.zstep;
        K.gather(tab, z1 + K.zstep, v[2]);
        if (!adj) K.gather(tab, z1, v[3]);
    }
}

__global__ void __launch_bounds__(QT_X * QP_WARPS, QP_MINB) sdf_query_grid_kernel(const __grid_constant__ QueryArgs A) {
    __shared__ MlpSmem sm;
    stage_mlp(A.f.mlp[0], sm);
    __syncthreads();
    const usl_grid_t &g = A.f.grid[0];
    const float2 *table = reinterpret_cast<const float2 *>(A.f.table[0]);
    constexpr int TZ = 2 * QP_WARPS;
    const int tx = (A.nx + QT_X - 1) / QT_X, tz = (A.nz + TZ - 1) / TZ;
    const int64_t tiles = (int64_t)(A.y_end - A.y_begin) * tx * tz;
    const int lx = threadIdx.x & (QT_X - 1), lz = threadIdx.x / QT_X;
    for (int64_t t = blockIdx.x; t < tiles; t += gridDim.x) {
        const int bz = (int)(t % tz);
        const int64_t r = t / tz;
        const int bx = (int)(r % tx);
        const int iy = (int)(r / tx) + A.y_begin;
        const int ix = bx * QT_X + lx, iz0 = bz * TZ + lz * 2;
        if (ix >= A.nx || iz0 >= A.nz) continue;
        auto norm = [&](float p, int d, bool &in) {                 // Mesher.py:147-156
            in = (p < A.f.bound_hi[d]) && (p > A.f.bound_lo[d]);
            const float x = __fdiv_rn(__fsub_rn(p, A.f.bound_lo[d]), __fsub_rn(A.f.bound_hi[d], A.f.bound_lo[d]));
            return fminf(fmaxf(x, 0.f), 1.f);
        };
        bool in_x, in_y, in_z0, in_z1;
        const float xn = norm(A.ax[ix], 0, in_x), yn = norm(A.ay[iy], 1, in_y);
        const float zn0 = norm(A.az[iz0], 2, in_z0), zn1 = norm(A.az[min(iz0 + 1, A.nz - 1)], 2, in_z1);

        f32x2_t hp0[USL_HID / 2], hp1[USL_HID / 2];
#pragma unroll
        for (int q = 0; q < USL_HID / 2; ++q) hp0[q] = hp1[q] = pack2(sm.b1[2 * q], sm.b1[2 * q + 1]);

        f32x2_t v[4][4];
        PairLevel nx_;
        auto issue = [&](int l) {
            const usl_level_t &lv = g.levels[l];
            const char *tab = reinterpret_cast<const char *>(table + lv.offset);
            uint32_t gx, gy, gz0, gz1;
            float wx, wy, wz0, wz1;
            pos_fract(lv.scale, xn, gx, wx); pos_fract(lv.scale, yn, gy, wy);
            pos_fract(lv.scale, zn0, gz0, wz0); pos_fract(lv.scale, zn1, gz1, wz1);
            nx_.same = gz1 == gz0; nx_.adj = gz1 == gz0 + 1u;
            nx_.wx2 = pack2(wx, wx); nx_.wy2 = pack2(wy, wy); nx_.wz0 = pack2(wz0, wz0); nx_.wz1 = pack2(wz1, wz1);
            if (lv.hashed) pair_issue<true>(lv, tab, gx, gy, gz0, gz1, nx_.same, nx_.adj, v);
            else pair_issue<false>(lv, tab, gx, gy, gz0, gz1, nx_.same, nx_.adj, v);
        };
        issue(0);
#pragma unroll 1
        for (int l = 0; l < g.n_levels; ++l) {
            // consume level l: collapse the planes in flight, pick each point's pair, lerp along z
            const PairLevel cur = nx_;
            const f32x2_t pa = plane_collapse(v[0], cur.wx2, cur.wy2), pb = plane_collapse(v[1], cur.wx2, cur.wy2);
            f32x2_t lo1 = pa, hi1 = pb;
            if (!cur.same) {
                hi1 = plane_collapse(v[2], cur.wx2, cur.wy2);
                lo1 = pb;
                if (!cur.adj) lo1 = plane_collapse(v[3], cur.wx2, cur.wy2);
            }
            const float2 f0 = unpack2(lerp2(cur.wz0, pa, pb)), f1 = unpack2(lerp2(cur.wz1, lo1, hi1));
            // level l + 1's gathers go out now and land while the 64 FMAs below issue
            if (l + 1 < g.n_levels) issue(l + 1);
            const f32x2_t fx0 = pack2(f0.x, f0.x), fy0 = pack2(f0.y, f0.y), fx1 = pack2(f1.x, f1.x), fy1 = pack2(f1.y, f1.y);
            const ulonglong2 *wa = reinterpret_cast<const ulonglong2 *>(sm.w1t[2 * l]);
            const ulonglong2 *wb = reinterpret_cast<const ulonglong2 *>(sm.w1t[2 * l + 1]);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const ulonglong2 a = wa[q], b = wb[q];
                ffma2(hp0[2 * q], a.x, fx0); ffma2(hp0[2 * q], b.x, fy0); ffma2(hp0[2 * q + 1], a.y, fx0); ffma2(hp0[2 * q + 1], b.y, fy0);
                ffma2(hp1[2 * q], a.x, fx1); ffma2(hp1[2 * q], b.x, fy1); ffma2(hp1[2 * q + 1], a.y, fx1); ffma2(hp1[2 * q + 1], b.y, fy1);
            }
        }
        float res[2];
#pragma unroll
        for (int p = 0; p < 2; ++p) {
            float h[USL_HID], th[1][USL_HID], out[4], tout[4][3];
#pragma unroll
            for (int q = 0; q < USL_HID / 2; ++q) { const float2 u = unpack2(p ? hp1[q] : hp0[q]); h[2 * q] = u.x; h[2 * q + 1] = u.y; }
            mlp_tail<false>(A.f.mlp[0], sm, h, th, out, tout);
            res[p] = (in_x && in_y && (p ? in_z1 : in_z0)) ? out[0] : -1.0f;     // Mesher.py:162
        }
        float *o = A.out + ((int64_t)(iy - A.y_begin) * A.nx + ix) * A.nz + iz0;   // (iy*nx + ix)*nz + iz, Mesher.py:192-193
        if ((A.nz & 1) == 0 && (reinterpret_cast<uintptr_t>(A.out) & 7) == 0) __stcs(reinterpret_cast<float2 *>(o), make_float2(res[0], res[1]));
        else { __stcs(o, res[0]); if (iz0 + 1 < A.nz) __stcs(o + 1, res[1]); }
    }
}

// ---- stand-alone decoder backward (tinycudann.Network / nn.Linear stacks: usl_mlp_bwd) -------------------
// feat is h[n,32] (row-major), raw / d_raw are [n,n_out]; gradients wrt the weights (block-reduced, one atomic per
// element per CTA) and wrt the input features dh[n,32].  The fused path has its own kernel (field_bwd.cu).
#define MB_THREADS 128
#define MB_WARPS (MB_THREADS / 32)
#define MB_STRIDE 52   // floats per point row: 16B-aligned rows, conflict-free 128-bit stores
struct MlpBwdArgs {
    usl_mlp_t m, gm;
    int has_gm;
    const float *h, *out, *dout;
    int64_t n;
    float *dh;
};

template <int NH>
__global__ void __launch_bounds__(MB_THREADS) mlp_bwd_kernel(const __grid_constant__ MlpBwdArgs A) {
    __shared__ MlpSmem sm;
    __shared__ __align__(16) float tiles[MB_WARPS][32][MB_STRIDE];
    const usl_mlp_t &m = A.m;
    stage_mlp(m, sm);
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float(*tile)[MB_STRIDE] = tiles[warp];
    const int64_t n = A.n;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = i < n;
    const float2 *fin = reinterpret_cast<const float2 *>(A.h) + (active ? i : 0) * (USL_IN / 2);
    float h1[USL_HID];
#pragma unroll
    for (int j = 0; j < USL_HID; ++j) h1[j] = sm.b1[j];
#pragma unroll
    for (int l = 0; l < USL_IN / 2; ++l) {
        float2 v = make_float2(0.f, 0.f);
        if (active) v = __ldg(fin + l);
        const float4 *wa = reinterpret_cast<const float4 *>(sm.w1t[2 * l]);
        const float4 *wb = reinterpret_cast<const float4 *>(sm.w1t[2 * l + 1]);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float4 a = wa[q], b = wb[q];
            h1[q * 4 + 0] = fmaf(b.x, v.y, fmaf(a.x, v.x, h1[q * 4 + 0]));
            h1[q * 4 + 1] = fmaf(b.y, v.y, fmaf(a.y, v.x, h1[q * 4 + 1]));
            h1[q * 4 + 2] = fmaf(b.z, v.y, fmaf(a.z, v.x, h1[q * 4 + 2]));
            h1[q * 4 + 3] = fmaf(b.w, v.y, fmaf(a.w, v.x, h1[q * 4 + 3]));
        }
    }
    float du[4] = {0.f, 0.f, 0.f, 0.f};
    if (active)
        for (int o = 0; o < m.n_out; ++o) du[o] = A.dout[i * m.n_out + o] * act_bwd(m.out_act, A.out[i * m.n_out + o]);
    float dh1[USL_HID];
    float a2[USL_HID], dh2[USL_HID];
    if (NH == 2) {
#pragma unroll
        for (int j = 0; j < USL_HID; ++j) dh1[j] = 0.f;
#pragma unroll
        for (int q = 0; q < USL_HID; ++q) {
            float s = sm.b2[q];
#pragma unroll
            for (int j = 0; j < USL_HID; ++j) s = fmaf(sm.w2[q][j], fmaxf(h1[j], 0.f), s);
            float d = 0.f;
#pragma unroll
            for (int o = 0; o < 4; ++o) d = fmaf(sm.wo[o][q], du[o], d);
            d = (s > 0.f) ? d : 0.f;
            a2[q] = fmaxf(s, 0.f);
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
    // weight gradients: warp-private tile, each lane owns a patch of every matrix
    float acc1[16], acc2[8], acco[2], accb[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int q = 0; q < 16; ++q) acc1[q] = 0.f;
#pragma unroll
    for (int q = 0; q < 8; ++q) acc2[q] = 0.f;
    acco[0] = acco[1] = 0.f;
    if (A.has_gm) {
        float4 *row = reinterpret_cast<float4 *>(tile[lane]);
#pragma unroll
        for (int q = 0; q < 4; ++q) row[q] = make_float4(dh1[4 * q], dh1[4 * q + 1], dh1[4 * q + 2], dh1[4 * q + 3]);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            float2 u0 = make_float2(0.f, 0.f), u1 = u0;
            if (active) { u0 = __ldg(fin + 2 * q); u1 = __ldg(fin + 2 * q + 1); }
            row[4 + q] = make_float4(u0.x, u0.y, u1.x, u1.y);
        }
        __syncwarp();
        {
            const int j = lane >> 1, k0 = (lane & 1) * 16;
#pragma unroll 4
            for (int p = 0; p < 32; ++p) {
                const float d = tile[p][j];
                const float4 *fr = reinterpret_cast<const float4 *>(&tile[p][16 + k0]);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float4 v = fr[q];
                    acc1[4 * q + 0] = fmaf(d, v.x, acc1[4 * q + 0]);
                    acc1[4 * q + 1] = fmaf(d, v.y, acc1[4 * q + 1]);
                    acc1[4 * q + 2] = fmaf(d, v.z, acc1[4 * q + 2]);
                    acc1[4 * q + 3] = fmaf(d, v.w, acc1[4 * q + 3]);
                }
                if (lane < 16) accb[0] += tile[p][lane];
            }
        }
        __syncwarp();
        // phase 2: [0:16] dh2, [16:32] a1, [32:36] du, [36:52] a_last
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            row[q] = (NH == 2) ? make_float4(dh2[4 * q], dh2[4 * q + 1], dh2[4 * q + 2], dh2[4 * q + 3]) : make_float4(0.f, 0.f, 0.f, 0.f);
            const float4 a1v = make_float4(fmaxf(h1[4 * q], 0.f), fmaxf(h1[4 * q + 1], 0.f), fmaxf(h1[4 * q + 2], 0.f), fmaxf(h1[4 * q + 3], 0.f));
            row[4 + q] = a1v;
            row[9 + q] = (NH == 2) ? make_float4(a2[4 * q], a2[4 * q + 1], a2[4 * q + 2], a2[4 * q + 3]) : a1v;
        }
        row[8] = make_float4(du[0], du[1], du[2], du[3]);
        __syncwarp();
        {
            const int r2 = lane >> 1, j0 = (lane & 1) * 8;
            const int o = lane >> 3, i0 = (lane & 7) * 2;
#pragma unroll 4
            for (int p = 0; p < 32; ++p) {
                if (NH == 2) {
                    const float d = tile[p][r2];
                    const float4 *ar = reinterpret_cast<const float4 *>(&tile[p][16 + j0]);
#pragma unroll
                    for (int q = 0; q < 2; ++q) {
                        const float4 v = ar[q];
                        acc2[4 * q + 0] = fmaf(d, v.x, acc2[4 * q + 0]);
                        acc2[4 * q + 1] = fmaf(d, v.y, acc2[4 * q + 1]);
                        acc2[4 * q + 2] = fmaf(d, v.z, acc2[4 * q + 2]);
                        acc2[4 * q + 3] = fmaf(d, v.w, acc2[4 * q + 3]);
                    }
                    if (lane < 16) accb[1] += tile[p][lane];
                }
                const float duo = tile[p][32 + o];
                const float2 al = *reinterpret_cast<const float2 *>(&tile[p][36 + i0]);
                acco[0] = fmaf(duo, al.x, acco[0]);
                acco[1] = fmaf(duo, al.y, acco[1]);
                if (lane < 4) accb[2] += tile[p][32 + lane];
            }
        }
        __syncwarp();
    }
    // gradient wrt the input features: d f_l = W1[:, 2l:2l+2]^T dh1
    if (active && A.dh != nullptr) {
#pragma unroll
        for (int l = 0; l < USL_IN / 2; ++l) {
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
            reinterpret_cast<float2 *>(A.dh)[i * (USL_IN / 2) + l] = make_float2(dfx, dfy);
        }
    }
    if (A.has_gm) {
        __syncthreads();
        float *red = &tiles[0][0][0];               // reuse: [MB_WARPS][32][33] floats needed (<= tile storage)
        float *mine = red + (warp * 32 + lane) * 33;
#pragma unroll
        for (int q = 0; q < 16; ++q) mine[q] = acc1[q];
#pragma unroll
        for (int q = 0; q < 8; ++q) mine[16 + q] = acc2[q];
        mine[24] = acco[0]; mine[25] = acco[1];
        mine[26] = accb[0]; mine[27] = accb[1]; mine[28] = accb[2];
        __syncthreads();
        const usl_mlp_t &gm = A.gm;
        for (int e = threadIdx.x; e < 32 * 29; e += blockDim.x) {
            const int ln = e / 29, q = e % 29;
            float s = 0.f;
#pragma unroll
            for (int w = 0; w < MB_WARPS; ++w) s += red[(w * 32 + ln) * 33 + q];
            if (q < 16) {                                   // dW1[j][k0+q]
                const int j = ln >> 1, k = (ln & 1) * 16 + q;
                if (gm.w1) atomicAdd(gm.w1 + j * USL_IN + k, s);
            } else if (q < 24) {                            // dW2[i][j0+q]
                if (NH == 2 && gm.w2) atomicAdd(gm.w2 + (ln >> 1) * USL_HID + (ln & 1) * 8 + (q - 16), s);
            } else if (q < 26) {                            // dWo[o][i0+q]
                const int o = ln >> 3, ii = (ln & 7) * 2 + (q - 24);
                if (o < m.n_out && gm.wo) atomicAdd(gm.wo + o * USL_HID + ii, s);
            } else if (q == 26) {
                if (ln < 16 && gm.b1) atomicAdd(gm.b1 + ln, s);
            } else if (q == 27) {
                if (NH == 2 && ln < 16 && gm.b2) atomicAdd(gm.b2 + ln, s);
            } else {
                if (ln < m.n_out && gm.bo) atomicAdd(gm.bo + ln, s);
            }
        }
    }
}

// ---- stand-alone decoder (tinycudann.Network / nn.Linear stacks on given features) -------------
__global__ void __launch_bounds__(256) mlp_fwd_kernel(usl_mlp_t m, const float *__restrict__ h, int64_t n,
                                                      float *__restrict__ out) {
    __shared__ MlpSmem sm;
    stage_mlp(m, sm);
    __syncthreads();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float a1[USL_HID];
#pragma unroll
    for (int j = 0; j < USL_HID; ++j) a1[j] = sm.b1[j];
    const float4 *hin = reinterpret_cast<const float4 *>(h + i * USL_IN);
#pragma unroll
    for (int q = 0; q < USL_IN / 4; ++q) {
        const float4 v = __ldg(hin + q);
        const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int e = 0; e < 4; ++e)
#pragma unroll
            for (int j = 0; j < USL_HID; ++j) a1[j] = fmaf(sm.w1t[q * 4 + e][j], vv[e], a1[j]);
    }
#pragma unroll
    for (int j = 0; j < USL_HID; ++j) a1[j] = fmaxf(a1[j], 0.f);
    float u[4] = {sm.bo[0], sm.bo[1], sm.bo[2], sm.bo[3]};
    if (m.n_hidden == 2) {
        for (int q = 0; q < USL_HID; ++q) {
            float s = sm.b2[q];
#pragma unroll
            for (int j = 0; j < USL_HID; ++j) s = fmaf(sm.w2[q][j], a1[j], s);
            s = fmaxf(s, 0.f);
#pragma unroll
            for (int o = 0; o < 4; ++o) u[o] = fmaf(sm.wo[o][q], s, u[o]);
        }
    } else {
#pragma unroll
        for (int o = 0; o < 4; ++o)
#pragma unroll
            for (int j = 0; j < USL_HID; ++j) u[o] = fmaf(sm.wo[o][j], a1[j], u[o]);
    }
    for (int o = 0; o < m.n_out; ++o) out[i * m.n_out + o] = act_fwd(m.out_act, u[o]);
}

static int check_field(const usl_field_t *f, const usl_points_t *p) {
    if (!f || !p) { set_error("null field/points"); return 1; }
    for (int gi = 0; gi < 2; ++gi) {
        if (f->grid[gi].n_levels != USL_IN / USL_FEATS) { set_error("field grids must have 16 levels x 2 features"); return 1; }
        if (f->mlp[gi].n_hidden < 1 || f->mlp[gi].n_hidden > 2 || f->mlp[gi].n_out < 1 || f->mlp[gi].n_out > 3) {
            set_error("unsupported decoder shape"); return 1;
        }
    }
    if (p->n > 0 && !p->x && (!p->rays_o || !p->rays_d || !p->z || p->S <= 0)) {     // an empty batch carries no pointers
        set_error("points: need x or (rays_o, rays_d, z, S)"); return 1;
    }
    return 0;
}

}  // namespace usl

using namespace usl;


extern "C" {

int usl_field_fwd(const usl_field_t *f, const usl_points_t *p, float *raw, float *feat, float *jac,
                  usl_stream_t stream) {
    if (check_field(f, p)) return 1;
    if (p->n <= 0) return 0;
    FieldArgs A;
    A.f = *f; A.p = *p; A.raw = raw; A.feat = feat; A.jac = jac; A.sdf = nullptr;
    dim3 grid((unsigned)((p->n + USL_FWD_THREADS - 1) / USL_FWD_THREADS), 2);
    cudaStream_t s = (cudaStream_t)stream;
    // USL_TCGEN05=1: tangent contraction of the Jacobian path on the tcgen05 tensor cores (field_tc.cu). Parity-tested
    // and profiled, but not the default: the kernel is bound by L1 sector lookups of the gather,
    // so moving 57 % of the FMA-pipe work to the tensor pipe does not shorten it (196 us vs 182 us measured).
    if (p->sample_major && !p->x && (p->n % p->S)) { set_error("usl_field_fwd: sample_major needs n == R * S"); return 1; }
    if (jac && feat) field_fwd_kernel<true, true><<<grid, USL_FWD_THREADS, 0, s>>>(A);
    else if (jac) field_fwd_kernel<true, false><<<grid, USL_FWD_THREADS, 0, s>>>(A);
    else if (feat) field_fwd_kernel<false, true><<<grid, USL_FWD_THREADS, 0, s>>>(A);
    else field_fwd_kernel<false, false><<<grid, USL_FWD_THREADS, 0, s>>>(A);
    return check_launch("usl_field_fwd");
}

int usl_field_sdf(const usl_field_t *f, const usl_points_t *p, float *sdf, usl_stream_t stream) {
    if (check_field(f, p)) return 1;
    if (p->n <= 0) return 0;
    if (p->sample_major) { set_error("usl_field_sdf: sample_major point order is only supported by usl_field_fwd"); return 1; }
    FieldArgs A;
    A.f = *f; A.p = *p; A.raw = nullptr; A.feat = nullptr; A.jac = nullptr; A.sdf = sdf;
    field_sdf_kernel<<<(unsigned)((p->n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(A);
    return check_launch("usl_field_sdf");
}

int usl_mlp_bwd(const usl_mlp_t *m, const usl_mlp_t *gm, const float *h, const float *out, const float *dout,
                int64_t n, float *dh, usl_stream_t stream) {
    if (!m || m->n_hidden < 1 || m->n_hidden > 2 || m->n_out < 1 || m->n_out > 3) { set_error("usl_mlp_bwd: unsupported decoder shape"); return 1; }
    if (n <= 0) return 0;
    MlpBwdArgs A;
    memset(&A, 0, sizeof(A));
    A.m = *m; A.has_gm = gm ? 1 : 0;
    if (gm) A.gm = *gm;
    A.h = h; A.out = out; A.dout = dout; A.n = n; A.dh = dh;
    const unsigned grid = (unsigned)((n + MB_THREADS - 1) / MB_THREADS);
    cudaStream_t s = (cudaStream_t)stream;
    if (m->n_hidden == 2) mlp_bwd_kernel<2><<<grid, MB_THREADS, 0, s>>>(A);
    else mlp_bwd_kernel<1><<<grid, MB_THREADS, 0, s>>>(A);
    return check_launch("usl_mlp_bwd");
}

int usl_mlp_fwd(const usl_mlp_t *m, const float *h, int64_t n, float *out, usl_stream_t stream) {
    if (!m || m->n_hidden < 1 || m->n_hidden > 2 || m->n_out < 1 || m->n_out > 3) { set_error("usl_mlp_fwd: unsupported decoder shape"); return 1; }
    if (n <= 0) return 0;
    mlp_fwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(*m, h, n, out);
    return check_launch("usl_mlp_fwd");
}

int usl_sdf_query_grid(const usl_field_t *f, const float *ax, const float *ay, const float *az, int nx, int ny,
                       int nz, int y_begin, int y_end, float *out, usl_stream_t stream) {
    if (!f || y_begin < 0 || y_end > ny || y_end < y_begin) { set_error("usl_sdf_query_grid: bad arguments"); return 1; }
    const int64_t total = (int64_t)(y_end - y_begin) * nx * nz;
    if (total <= 0) return 0;
    QueryArgs A;
    A.f = *f; A.ax = ax; A.ay = ay; A.az = az; A.nx = nx; A.ny = ny; A.nz = nz; A.y_begin = y_begin; A.y_end = y_end; A.out = out;
    int64_t blocks = (int64_t)(y_end - y_begin) * ((nx + QT_X - 1) / QT_X) * ((nz + 2 * QP_WARPS - 1) / (2 * QP_WARPS));   // one tile per CTA and pass
    int dev = 0, n_sm = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    const int64_t cap = (int64_t)(n_sm > 0 ? n_sm : 1) * 64;     // persistent tile loop: a few waves of resident CTAs per SM
    if (blocks > cap) blocks = cap;
    sdf_query_grid_kernel<<<(unsigned)blocks, QT_X * QP_WARPS, 0, (cudaStream_t)stream>>>(A);
    return check_launch("usl_sdf_query_grid");
}

}  // extern "C"
