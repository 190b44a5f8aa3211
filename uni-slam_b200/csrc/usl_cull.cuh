// Element functions of the mesh culling kernels (cull.cu): the arithmetic of src/tools/cull_mesh.py for ONE vertex against
// ONE frame / one hull plane / one face.  Written without device-only constructs so that the same source also compiles with
// plain g++: tests/host_harness/cull_host.cpp runs these functions over a whole mesh on the host, which is how the arithmetic
// is checked in a container without a GPU.  The harness is test infrastructure; the library never runs this code on the CPU.
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

// cull_mesh.py:69-99 for one vertex and one frame.  w2c: the 12 floats of the first three rows of torch.inverse(c2w);
// depth: the frame's (H,W) sensor depth, read only with eval_rec and only for vertices inside the frustum.
USL_HD bool cull_seen_in_frame(float px, float py, float pz, const float *w2c, const float *depth, const CullCam &c) {
    // cam = (w2c @ [p, 1])[:3];  cam[0] *= -1
    float cam[3];
    for (int r = 0; r < 3; ++r) {
        const float *m = w2c + 4 * r;
        cam[r] = USLC_ADD(USLC_ADD(USLC_ADD(USLC_MUL(m[0], px), USLC_MUL(m[1], py)), USLC_MUL(m[2], pz)), m[3]);
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
    int x0 = (int)x0f, y0 = (int)y0f;
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

// mesh_bound.contains for a closed convex hull (cull_mesh.py:136-142): signed distance to one outward plane (n, d)
USL_HD float cull_plane_side(float px, float py, float pz, const float *plane) {
    return USLC_ADD(USLC_ADD(USLC_ADD(USLC_MUL(plane[0], px), USLC_MUL(plane[1], py)), USLC_MUL(plane[2], pz)), plane[3]);
}

// cull_mesh.py:101-102 (require_all = 0: a face goes when all three vertices are unseen, i.e. stays when any is seen) and
// :143-144 (require_all = 1: a face stays when all three vertices are inside the bound)
USL_HD bool cull_face_keep(uint8_t m0, uint8_t m1, uint8_t m2, int require_all) {
    return require_all ? (m0 && m1 && m2) : (m0 || m1 || m2);
}

}  // namespace usl
