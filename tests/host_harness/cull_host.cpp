// TEST INFRASTRUCTURE, NOT PRODUCT CODE.
// Runs the thread functions of uni-slam_b200/csrc/usl_cull.cuh -- the very source the CUDA kernels of cull.cu are trampolines
// for -- for every thread of a simulated grid on the host: `nthreads` threads along x (the kernels' gridDim.x * 256), frame
// groups in blockIdx.y order.  Built by loader.py with g++ (-ffp-contract=off) into tests/_build/; loaded only by tests and by
// bench.py's isolated culling leg as the checker.  It exists because the build container has no GPU: the kernels' per-thread
// code is executed here against the oracle and the reference-generated golden, and the -m gpu tests then require the kernels to
// reproduce this harness bit for bit.  The product library has no CPU path.
#include <stdint.h>
#include <string.h>

#include "../../uni-slam_b200/csrc/usl_cull.cuh"

extern "C" {

void cull_host_frames(const float *verts, int64_t V, const float *w2c, const float *depths, int K, int H, int W, float fx, float fy,
                      float cx, float cy, float truncation, int eval_rec, int frames_per_cta, int64_t nthreads, uint8_t *seen) {
    usl::CullFramesArgs A;
    A.verts = verts; A.V = V; A.w2c = w2c; A.depths = eval_rec ? depths : nullptr; A.K = K; A.frames_per_cta = frames_per_cta;
    A.cam = usl::CullCam{H, W, fx, fy, cx, cy, truncation, eval_rec ? 1 : 0};
    A.seen = seen;
    const int groups = (K + frames_per_cta - 1) / frames_per_cta;
    for (int g = 0; g < groups; ++g)
        for (int64_t tid = 0; tid < nthreads; ++tid) usl::cull_frames_thread(A, tid, nthreads, g);
}

void cull_host_hull(const float *verts, int64_t V, const float *planes, int F, int64_t nthreads, uint8_t *inside) {
    for (int64_t tid = 0; tid < nthreads; ++tid) usl::cull_hull_thread(verts, V, planes, F, inside, tid, nthreads);
}

// mesh_face_keep_kernel + exclusive scans (stand-ins for usl_scan_u8) + mesh_compact_kernel; the new counts come back through
// n_out[0] (vertices), n_out[1] (faces)
void cull_host_compact(const float *verts, const uint8_t *colors, int64_t V, const int32_t *faces, int64_t T, const uint8_t *vmask,
                       int require_all, int64_t nthreads, uint8_t *keep, float *verts_out, uint8_t *colors_out, int32_t *faces_out,
                       int64_t *n_out) {
    uint8_t *vref = new uint8_t[V > 0 ? V : 1]();
    uint32_t *voff = new uint32_t[V > 0 ? V : 1];
    uint32_t *foff = new uint32_t[T > 0 ? T : 1];
    for (int64_t tid = 0; tid < nthreads; ++tid) usl::cull_face_keep_thread(faces, T, vmask, V, require_all, keep, vref, tid, nthreads);
    uint32_t nv = 0, nf = 0;
    for (int64_t v = 0; v < V; ++v) { voff[v] = nv; nv += vref[v]; }
    for (int64_t t = 0; t < T; ++t) { foff[t] = nf; nf += keep[t]; }
    usl::CompactArgs A{verts, colors, V, faces, T, keep, vref, voff, foff, verts_out, colors_out, faces_out};
    for (int64_t tid = 0; tid < nthreads; ++tid) usl::cull_compact_thread(A, tid, nthreads);
    n_out[0] = nv; n_out[1] = nf;
    delete[] vref; delete[] voff; delete[] foff;
}

}  // extern "C"
