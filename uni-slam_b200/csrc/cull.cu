// f4, the culling after marching cubes: src/tools/cull_mesh.py on the device-resident mesh.
//
//   usl_mesh_cull_frames   cull_mesh (cull_mesh.py:58-99): a vertex is "seen" when some frame has it inside its frustum, in
//                          front of the camera and (eval_rec) not behind the sensor depth + truncation.  The reference walks
//                          the frames on the host: per frame ~20 torch launches over all vertices, a grid_sample, a D2H copy of
//                          the mask and a numpy AND.  Here: thread = vertex, blockIdx.y = a group of consecutive frames; a vertex
//                          leaves at the first frame that sees it, and a vertex that an earlier group already marked is skipped
//                          at entry, so the work is dominated by the vertices no frame sees.  Frame groups are the slow grid
//                          dimension: the CTAs in flight work on the same few depth frames (16 x 3.3 MB at 1200x680), which
//                          therefore stay in L2 while all vertex tiles pass over them; with 180 GB of HBM the whole sequence's
//                          depth frames (2000 x 3.3 MB = 6.5 GB) stay resident.  The camera rows are read at warp-uniform
//                          addresses (one broadcast transaction per load, L1-resident).
//   usl_mesh_cull_hull     cull_out_bound_mesh (cull_mesh.py:137-143): vertex inside the closed convex bound = on the inner side
//                          of every hull plane (warp-uniform plane reads, a thread leaves at the first plane it is outside of)
//   usl_mesh_face_keep     the face rule of either culling (cull_mesh.py:100-101 / :144-145) + the referenced-vertex marks
//   usl_mesh_compact       update_faces + remove_unreferenced_vertices (order-preserving) from the exclusive scans of the keep
//                          flags and the vertex marks (usl_scan_u8 of mesh.cu)
//
// The kernels are trampolines: what a thread does lives in usl_cull.cuh, which also compiles for the host -- the test harness
// (tests/host_harness) runs those thread functions over a simulated grid, and the -m gpu tests hold the kernels to it bit for bit.
#include "usl_device.cuh"
#include "usl_cull.cuh"

namespace usl {

constexpr int CULL_THREADS = 256;
constexpr int CULL_MAX_FRAMES_PER_CTA = 64;

#define USL_CULL_TID ((int64_t)blockIdx.x * CULL_THREADS + threadIdx.x)
#define USL_CULL_NTHREADS ((int64_t)gridDim.x * CULL_THREADS)

__global__ void __launch_bounds__(CULL_THREADS) mesh_cull_frames_kernel(const __grid_constant__ CullFramesArgs A) {
    cull_frames_thread(A, USL_CULL_TID, USL_CULL_NTHREADS, (int)blockIdx.y);
}

__global__ void __launch_bounds__(CULL_THREADS) mesh_cull_hull_kernel(const float *__restrict__ verts, int64_t V, const float *__restrict__ planes,
                                                                      int F, uint8_t *__restrict__ inside) {
    cull_hull_thread(verts, V, planes, F, inside, USL_CULL_TID, USL_CULL_NTHREADS);
}

__global__ void __launch_bounds__(CULL_THREADS) mesh_face_keep_kernel(const int32_t *__restrict__ faces, int64_t T, const uint8_t *__restrict__ vmask,
                                                                      int64_t V, int require_all, uint8_t *__restrict__ keep,
                                                                      uint8_t *__restrict__ vref) {
    cull_face_keep_thread(faces, T, vmask, V, require_all, keep, vref, USL_CULL_TID, USL_CULL_NTHREADS);
}

__global__ void __launch_bounds__(CULL_THREADS) mesh_compact_kernel(const __grid_constant__ CompactArgs A) {
    cull_compact_thread(A, USL_CULL_TID, USL_CULL_NTHREADS);
}

// grid-stride launches: every tile of 256 items, capped at ctas_per_sm CTAs per SM of this device
static unsigned cull_grid(int64_t n, int ctas_per_sm) {
    int dev = 0, n_sm = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    const int64_t b = (n + CULL_THREADS - 1) / CULL_THREADS;
    const int64_t cap = (int64_t)(n_sm > 0 ? n_sm : 1) * ctas_per_sm;
    return (unsigned)(b < cap ? (b < 1 ? 1 : b) : cap);
}

}  // namespace usl

using namespace usl;

extern "C" {

int usl_mesh_cull_frames(const usl_cull_frames_args_t *a, usl_stream_t stream) {
    if (!a) { set_error("usl_mesh_cull_frames: null argument"); return 1; }
    if (a->V <= 0 || a->K <= 0) return 0;
    if (!a->verts || !a->w2c || !a->seen || a->H < 2 || a->W < 2 || (a->eval_rec && !a->depths)) {
        set_error("usl_mesh_cull_frames: bad arguments (verts, w2c, seen required; depths required with eval_rec; H, W >= 2)"); return 1;
    }
    int fpc = a->frames_per_cta > 0 ? a->frames_per_cta : 16;
    if (fpc > CULL_MAX_FRAMES_PER_CTA) fpc = CULL_MAX_FRAMES_PER_CTA;
    const int groups = (a->K + fpc - 1) / fpc;
    if (groups > 65535) { set_error("usl_mesh_cull_frames: %d frames in groups of %d exceed the grid's y extent; call per range of frames", a->K, fpc); return 1; }
    CullFramesArgs A;
    A.verts = a->verts; A.V = a->V; A.w2c = a->w2c; A.depths = a->eval_rec ? a->depths : nullptr; A.K = a->K; A.frames_per_cta = fpc;
    A.cam.H = a->H; A.cam.W = a->W; A.cam.fx = a->fx; A.cam.fy = a->fy; A.cam.cx = a->cx; A.cam.cy = a->cy;
    A.cam.truncation = a->truncation; A.cam.eval_rec = a->eval_rec ? 1 : 0;
    A.seen = a->seen;
    // x extent capped at 32 CTAs per SM worth of vertex tiles (grid-stride beyond), so that one frame group's CTAs drain before
    // the next group's depth frames are pulled into L2
    dim3 grid(cull_grid(a->V, 32), (unsigned)groups);
    mesh_cull_frames_kernel<<<grid, CULL_THREADS, 0, (cudaStream_t)stream>>>(A);
    return check_launch("usl_mesh_cull_frames");
}

int usl_mesh_cull_hull(const float *verts, int64_t V, const float *planes, int32_t F, uint8_t *inside, usl_stream_t stream) {
    if (V <= 0) return 0;
    if (!verts || !inside || F < 0 || (F > 0 && !planes)) { set_error("usl_mesh_cull_hull: bad arguments"); return 1; }
    mesh_cull_hull_kernel<<<cull_grid(V, 32), CULL_THREADS, 0, (cudaStream_t)stream>>>(verts, V, planes, F, inside);
    return check_launch("usl_mesh_cull_hull");
}

int usl_mesh_face_keep(const int32_t *faces, int64_t T, const uint8_t *vmask, int64_t V, int32_t require_all, uint8_t *keep,
                       uint8_t *vref, usl_stream_t stream) {
    if (T <= 0) return 0;
    if (!faces || !vmask || !keep || !vref || V < 0) { set_error("usl_mesh_face_keep: bad arguments"); return 1; }
    mesh_face_keep_kernel<<<cull_grid(T, 16), CULL_THREADS, 0, (cudaStream_t)stream>>>(faces, T, vmask, V, require_all ? 1 : 0, keep, vref);
    return check_launch("usl_mesh_face_keep");
}

int usl_mesh_compact(const float *verts, const uint8_t *colors, int64_t V, const int32_t *faces, int64_t T, const uint8_t *keep,
                     const uint8_t *vref, const uint32_t *voff, const uint32_t *foff, float *verts_out, uint8_t *colors_out,
                     int32_t *faces_out, usl_stream_t stream) {
    if (V <= 0 && T <= 0) return 0;
    if (!verts || !vref || !voff || !verts_out || (T > 0 && (!faces || !keep || !foff || !faces_out)) || ((colors != nullptr) != (colors_out != nullptr))) {
        set_error("usl_mesh_compact: bad arguments"); return 1;
    }
    CompactArgs A{verts, colors, V, faces, T, keep, vref, voff, foff, verts_out, colors_out, faces_out};
    mesh_compact_kernel<<<cull_grid(V > T ? V : T, 16), CULL_THREADS, 0, (cudaStream_t)stream>>>(A);
    return check_launch("usl_mesh_compact");
}

}  // extern "C"
