// TEST INFRASTRUCTURE, NOT PRODUCT CODE.
// Runs the element functions of uni-slam_b200/csrc/usl_cull.cuh -- the very source the CUDA kernels of cull.cu inline -- over a
// whole mesh on the host, with the loop structure of the kernels written out serially.  Built by tests/helpers.py with g++
// (-ffp-contract=off) into tests/_build/; loaded only by tests.  It exists because the build container has no GPU: the
// arithmetic of the culling kernels is checked here against the oracle and the reference-generated golden, and the -m gpu
// tests then require the kernels to reproduce this harness bit for bit.  The product library has no CPU path.
#include <stdint.h>
#include <string.h>

#include "../../uni-slam_b200/csrc/usl_cull.cuh"

extern "C" {

// the frame loop of mesh_cull_frames_kernel: groups of frames_per_cta frames, early exit per vertex, OR-accumulated marks
void cull_host_frames(const float *verts, int64_t V, const float *w2c, const float *depths, int K, int H, int W, float fx, float fy,
                      float cx, float cy, float truncation, int eval_rec, int frames_per_cta, uint8_t *seen) {
    usl::CullCam cam{H, W, fx, fy, cx, cy, truncation, eval_rec ? 1 : 0};
    const int64_t frame_px = (int64_t)H * W;
    for (int k0 = 0; k0 < K; k0 += frames_per_cta) {
        const int nk = (K - k0 < frames_per_cta) ? K - k0 : frames_per_cta;
        for (int64_t v = 0; v < V; ++v) {
            if (seen[v]) continue;
            bool s = false;
            for (int k = 0; k < nk && !s; ++k) {
                float rows[12];
                memcpy(rows, w2c + (int64_t)(k0 + k) * 16, sizeof(rows));
                s = usl::cull_seen_in_frame(verts[v * 3], verts[v * 3 + 1], verts[v * 3 + 2], rows,
                                            (eval_rec && depths) ? depths + (int64_t)(k0 + k) * frame_px : nullptr, cam);
            }
            if (s) seen[v] = 1;
        }
    }
}

void cull_host_hull(const float *verts, int64_t V, const float *planes, int F, uint8_t *inside) {
    for (int64_t v = 0; v < V; ++v) {
        bool in = true;
        for (int f = 0; f < F; ++f) in = in && (usl::cull_plane_side(verts[v * 3], verts[v * 3 + 1], verts[v * 3 + 2], planes + f * 4) <= 0.f);
        inside[v] = in ? 1 : 0;
    }
}

// mesh_face_keep_kernel + exclusive scans + mesh_compact_kernel; returns the new counts through n_out[0] (vertices), n_out[1] (faces)
void cull_host_compact(const float *verts, const uint8_t *colors, int64_t V, const int32_t *faces, int64_t T, const uint8_t *vmask,
                       int require_all, uint8_t *keep, float *verts_out, uint8_t *colors_out, int32_t *faces_out, int64_t *n_out) {
    uint8_t *vref = new uint8_t[V > 0 ? V : 1]();
    uint32_t *voff = new uint32_t[V > 0 ? V : 1];
    for (int64_t t = 0; t < T; ++t) {
        const int32_t a = faces[t * 3], b = faces[t * 3 + 1], c = faces[t * 3 + 2];
        bool k = false;
        if (a >= 0 && a < V && b >= 0 && b < V && c >= 0 && c < V) {
            k = usl::cull_face_keep(vmask[a], vmask[b], vmask[c], require_all);
            if (k) { vref[a] = 1; vref[b] = 1; vref[c] = 1; }
        }
        keep[t] = k ? 1 : 0;
    }
    uint32_t nv = 0;
    for (int64_t v = 0; v < V; ++v) { voff[v] = nv; nv += vref[v]; }
    for (int64_t v = 0; v < V; ++v) {
        if (!vref[v]) continue;
        for (int d = 0; d < 3; ++d) verts_out[(int64_t)voff[v] * 3 + d] = verts[v * 3 + d];
        if (colors && colors_out)
            for (int d = 0; d < 3; ++d) colors_out[(int64_t)voff[v] * 3 + d] = colors[v * 3 + d];
    }
    int64_t nf = 0;
    for (int64_t t = 0; t < T; ++t) {
        if (!keep[t]) continue;
        for (int d = 0; d < 3; ++d) faces_out[nf * 3 + d] = (int32_t)voff[faces[t * 3 + d]];
        ++nf;
    }
    n_out[0] = nv; n_out[1] = nf;
    delete[] vref; delete[] voff;
}

}  // extern "C"
