// Thread functions of the mesh culling kernels (cull.cu): everything one CUDA thread of those kernels does -- its grid-stride
// loop, the arithmetic of src/tools/cull_mesh.py for a vertex against a frame / a hull plane, the face rule, the compaction
// writes.  The kernels are trampolines that pass their thread index in.  Written without device-only constructs so that the
// same source also compiles with plain g++: tests/host_harness/cull_host.cpp runs these functions for every thread of a
// simulated grid on the host, which is how the kernels are checked in a container without a GPU.  The harness is test
// infrastructure; the library never runs this code on the CPU.
//
// Every product and sum is rounded on its own (no contraction into FMAs), so host and device agree bit for bit.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define USL_HD __host__ __device__ __forceinline__
#else
#define USL_HD static inline
#endif

#if defined(__CUDA_ARCH__)
#define USLC_MUL(a, b) __fmul_rn((a), (b))
#define USLC_ADD(a, b) __fadd_rn((a), (b))
#define USLC_SUB(a, b) __fsub_rn((a), (b))
#define USLC_DIV(a, b) __fdiv_rn((a), (b))
#define USLC_LD(p) __ldg(p)
#else
#define USLC_MUL(a, b) ((a) * (b))
#define USLC_ADD(a, b) ((a) + (b))
#define USLC_SUB(a, b) ((a) - (b))
#define USLC_DIV(a, b) ((a) / (b))
#define USLC_LD(p) (*(p))
#endif

namespace usl {

struct CullCam {
    int H, W;
    float fx, fy, cx, cy;
    float truncation;
    int eval_rec;
};

// cull_mesh.py:69-99 for one vertex and one frame.  w2c: the frame's torch.inverse(c2w), row-major (the first three rows are
// read; every thread of a warp reads the same 12 addresses: one broadcast transaction each); depth: the frame's (H,W) sensor
// depth, read only with eval_rec and only for vertices inside the frustum.
USL_HD bool cull_seen_in_frame(float px, float py, float pz, const float *w2c, const float *depth, const CullCam &c) {
    // cam = (w2c @ [p, 1])[:3];  cam[0] *= -1
    float cam[3];
    for (int r = 0; r < 3; ++r) {
        const float *m = w2c + 4 * r;
        cam[r] = USLC_ADD(USLC_ADD(USLC_ADD(USLC_MUL(USLC_LD(m), px), USLC_MUL(USLC_LD(m + 1), py)), USLC_MUL(USLC_LD(m + 2), pz)), USLC_LD(m + 3));
    }
    cam[0] = -cam[0];
    // uv = K @ cam;  z = uv[2] + 1e-5;  uv = uv[:2] / z
    const float z = USLC_ADD(cam[2], 1e-5f);
    const float u = USLC_DIV(USLC_ADD(USLC_MUL(c.fx, cam[0]), USLC_MUL(c.cx, cam[2])), z);
    const float v = USLC_DIV(USLC_ADD(USLC_MUL(c.fy, cam[1]), USLC_MUL(c.cy, cam[2])), z);
    const float fW = (float)c.W, fH = (float)c.H;
    // (0 <= -z) & (u < W) & (u > 0) & (v < H) & (v > 0)   (edge = 0); NaN / inf coordinates fail the comparisons
    if (!((0.f <= -z) && (u < fW) && (u > 0.f) && (v < fH) && (v > 0.f))) return false;
    if (!c.eval_rec) return true;
    // depth_samples = grid_sample(depth, 2 * (u / W, v / H) - 1, bilinear, align_corners=True, zeros):
    // source index = (g + 1) * (size - 1) / 2; inside the frustum 0 < u < W, so the four taps are inside the image
    const float gx = USLC_SUB(USLC_MUL(2.f, USLC_DIV(u, fW)), 1.f), gy = USLC_SUB(USLC_MUL(2.f, USLC_DIV(v, fH)), 1.f);
    const float ix = USLC_MUL(USLC_ADD(gx, 1.f), (float)(c.W - 1) / 2.f), iy = USLC_MUL(USLC_ADD(gy, 1.f), (float)(c.H - 1) / 2.f);
    const float x0f = floorf(ix), y0f = floorf(iy);
    const int x0 = (int)x0f, y0 = (int)y0f;
    const float w = USLC_SUB(ix, x0f), e = USLC_SUB(1.f, w), n = USLC_SUB(iy, y0f), s = USLC_SUB(1.f, n);
    float ds = 0.f;
    // zero padding: a tap outside the image contributes nothing (reachable only through rounding at the far border)
    if (y0 >= 0 && y0 < c.H) {
        const float *row = depth + (int64_t)y0 * c.W;
        if (x0 >= 0 && x0 < c.W) ds = USLC_ADD(ds, USLC_MUL(USLC_LD(row + x0), USLC_MUL(e, s)));
        if (x0 + 1 >= 0 && x0 + 1 < c.W) ds = USLC_ADD(ds, USLC_MUL(USLC_LD(row + x0 + 1), USLC_MUL(w, s)));
    }
    if (y0 + 1 >= 0 && y0 + 1 < c.H) {
        const float *row = depth + (int64_t)(y0 + 1) * c.W;
        if (x0 >= 0 && x0 < c.W) ds = USLC_ADD(ds, USLC_MUL(USLC_LD(row + x0), USLC_MUL(e, n)));
        if (x0 + 1 >= 0 && x0 + 1 < c.W) ds = USLC_ADD(ds, USLC_MUL(USLC_LD(row + x0 + 1), USLC_MUL(w, n)));
    }
    // depth_samples + truncation >= -z
    return USLC_ADD(ds, c.truncation) >= -z;
}

struct CullFramesArgs {
    const float *verts;       // [V,3]
    int64_t V;
    const float *w2c;         // [K,4,4]
    const float *depths;      // [K,H,W] or null (eval_rec == 0)
    int K, frames_per_cta;
    CullCam cam;
    uint8_t *seen;            // [V] OR-accumulated
};

// One thread of mesh_cull_frames_kernel: thread `tid` of `nthreads` along the grid's x extent, frame group `group` (= blockIdx.y).
// A vertex that an earlier group (or an earlier call) already marked is skipped; a vertex leaves at the first frame that sees it.
USL_HD void cull_frames_thread(const CullFramesArgs &A, int64_t tid, int64_t nthreads, int group) {
    const int k0 = group * A.frames_per_cta;
    const int nk = (A.K - k0 < A.frames_per_cta) ? (A.K - k0) : A.frames_per_cta;
    const int64_t frame_px = (int64_t)A.cam.H * A.cam.W;
    for (int64_t v = tid; v < A.V; v += nthreads) {
        if (A.seen[v]) continue;
        const float px = A.verts[v * 3], py = A.verts[v * 3 + 1], pz = A.verts[v * 3 + 2];
        bool s = false;
        for (int k = 0; k < nk && !s; ++k)
            s = cull_seen_in_frame(px, py, pz, A.w2c + (int64_t)(k0 + k) * 16, A.depths ? A.depths + (int64_t)(k0 + k) * frame_px : nullptr, A.cam);
        if (s) A.seen[v] = 1;                             // every writer stores the same value
    }
}

// One thread of mesh_cull_hull_kernel.  mesh_bound.contains for a closed convex hull (cull_mesh.py:137-143): inside = on the
// inner side (n . p + d <= 0) of every outward plane; planes[F,4] are read at warp-uniform addresses.
USL_HD void cull_hull_thread(const float *verts, int64_t V, const float *planes, int F, uint8_t *inside, int64_t tid, int64_t nthreads) {
    for (int64_t v = tid; v < V; v += nthreads) {
        const float px = verts[v * 3], py = verts[v * 3 + 1], pz = verts[v * 3 + 2];
        bool in = true;
        for (int f = 0; f < F && in; ++f) {
            const float *pl = planes + (int64_t)f * 4;
            const float side = USLC_ADD(USLC_ADD(USLC_ADD(USLC_MUL(USLC_LD(pl), px), USLC_MUL(USLC_LD(pl + 1), py)), USLC_MUL(USLC_LD(pl + 2), pz)), USLC_LD(pl + 3));
            in = side <= 0.f;
        }
        inside[v] = in ? 1 : 0;
    }
}

// One thread of mesh_face_keep_kernel.  cull_mesh.py:100-101 (require_all = 0: a face goes when all three vertices are unseen,
// i.e. stays when any is seen) and :144-145 (require_all = 1: a face stays when all three vertices are inside the bound);
// the vertices of a kept face are marked referenced.  A face with an index outside the vertex array is dropped.
USL_HD void cull_face_keep_thread(const int32_t *faces, int64_t T, const uint8_t *vmask, int64_t V, int require_all, uint8_t *keep,
                                  uint8_t *vref, int64_t tid, int64_t nthreads) {
    for (int64_t t = tid; t < T; t += nthreads) {
        const int32_t a = faces[t * 3], b = faces[t * 3 + 1], c = faces[t * 3 + 2];
        bool k = false;
        if (a >= 0 && a < V && b >= 0 && b < V && c >= 0 && c < V) {
            const uint8_t m0 = vmask[a], m1 = vmask[b], m2 = vmask[c];
            k = require_all ? (m0 && m1 && m2) : (m0 || m1 || m2);
            if (k) { vref[a] = 1; vref[b] = 1; vref[c] = 1; }
        }
        keep[t] = k ? 1 : 0;
    }
}

struct CompactArgs {
    const float *verts;       // [V,3]
    const uint8_t *colors;    // [V,3] or null
    int64_t V;
    const int32_t *faces;     // [T,3]
    int64_t T;
    const uint8_t *keep, *vref;
    const uint32_t *voff, *foff;      // exclusive scans of vref / keep
    float *verts_out;
    uint8_t *colors_out;
    int32_t *faces_out;
};

// One thread of mesh_compact_kernel: trimesh's update_faces + remove_unreferenced_vertices, order-preserving.
USL_HD void cull_compact_thread(const CompactArgs &A, int64_t tid, int64_t nthreads) {
    for (int64_t v = tid; v < A.V; v += nthreads) {
        if (!A.vref[v]) continue;
        const int64_t o = A.voff[v];
        for (int d = 0; d < 3; ++d) A.verts_out[o * 3 + d] = A.verts[v * 3 + d];
        if (A.colors && A.colors_out)
            for (int d = 0; d < 3; ++d) A.colors_out[o * 3 + d] = A.colors[v * 3 + d];
    }
    for (int64_t t = tid; t < A.T; t += nthreads) {
        if (!A.keep[t]) continue;
        const int64_t o = A.foff[t];
        for (int d = 0; d < 3; ++d) A.faces_out[o * 3 + d] = (int32_t)A.voff[A.faces[t * 3 + d]];
    }
}

}  // namespace usl
