// TEST INFRASTRUCTURE, NOT PRODUCT CODE.
// A torch-free consumer of the C-ABI (include/unislam_b200.h + libunislam_b200.so, plain cudaMalloc'd pointers): runs the mesh
// culling entry points and usl_render_metrics on a GPU and compares them with the host build of the very thread functions the
// kernels are trampolines for (uni-slam_b200/csrc/usl_cull.cuh, usl_metrics.cuh).  Starts in about a second (no Python, no
// torch), prints one JSON object.  Exit code 0 = every comparison exact (metrics: 1e-12), 1 = mismatch, 2 = CUDA / ABI error.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -cudart shared -o tests/_build/cull_gpu_check tests/host_harness/cull_gpu_check.cu \
//        -Iinclude -Luni-slam_b200/lib -lunislam_b200 -Xlinker -rpath -Xlinker '$ORIGIN/../../uni-slam_b200/lib'
//   tests/_build/cull_gpu_check [n_vertices] [n_frames]
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../../include/unislam_b200.h"
#include "../../uni-slam_b200/csrc/usl_cull.cuh"
#include "../../uni-slam_b200/csrc/usl_metrics.cuh"

#define CK(x)                                                                                       \
    do {                                                                                            \
        cudaError_t e_ = (x);                                                                       \
        if (e_ != cudaSuccess) { printf("{\"error\": \"%s: %s\"}\n", #x, cudaGetErrorString(e_)); return 2; } \
    } while (0)
#define ABI(x)                                                                                      \
    do {                                                                                            \
        if ((x) != 0) { printf("{\"error\": \"%s: %s\"}\n", #x, usl_last_error()); return 2; }      \
    } while (0)

static uint32_t g_state = 12345u;
static float frand() {                      // platform-independent LCG in [0,1)
    g_state = g_state * 1664525u + 1013904223u;
    return (float)(g_state >> 8) * (1.0f / 16777216.0f);
}

template <class T>
static T *to_device(const std::vector<T> &h) {
    T *d = nullptr;
    if (cudaMalloc(&d, h.size() * sizeof(T) + 16) != cudaSuccess) return nullptr;
    cudaMemcpy(d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice);
    return d;
}

int main(int argc, char **argv) {
    const int64_t V = argc > 1 ? atoll(argv[1]) : 1000000;
    const int K = argc > 2 ? atoi(argv[2]) : 64;
    const int H = 680, W = 1200;
    const float fx = 600.f, fy = 600.f, cx = 599.5f, cy = 339.5f, tr = 0.06f;
    // vertices in a 6 x 4 x 3 box; cameras on a circle inside it looking outwards (OpenGL frame: -z forward, +y up);
    // w2c is written directly: rows = camera axes, translation = -R^T c
    std::vector<float> verts(V * 3), w2c((size_t)K * 16), depths((size_t)K * H * W);
    for (int64_t v = 0; v < V; ++v) { verts[v * 3] = 6.f * frand() - 3.f; verts[v * 3 + 1] = 4.f * frand() - 2.f; verts[v * 3 + 2] = 3.f * frand() - 1.5f; }
    for (int k = 0; k < K; ++k) {
        const float a = 6.2831853f * k / K, c[3] = {0.8f * cosf(a), 0.6f * sinf(a), 0.1f * sinf(3 * a)};
        const float fwd[3] = {cosf(a + 0.3f), sinf(a + 0.3f), 0.f};
        const float zc[3] = {-fwd[0], -fwd[1], -fwd[2]}, up[3] = {0, 0, 1};
        float xc[3] = {up[1] * zc[2] - up[2] * zc[1], up[2] * zc[0] - up[0] * zc[2], up[0] * zc[1] - up[1] * zc[0]};
        const float xn = sqrtf(xc[0] * xc[0] + xc[1] * xc[1] + xc[2] * xc[2]);
        for (int d = 0; d < 3; ++d) xc[d] /= xn;
        const float yc[3] = {zc[1] * xc[2] - zc[2] * xc[1], zc[2] * xc[0] - zc[0] * xc[2], zc[0] * xc[1] - zc[1] * xc[0]};
        const float *ax[3] = {xc, yc, zc};
        float *m = &w2c[(size_t)k * 16];
        for (int r = 0; r < 3; ++r) {
            for (int d = 0; d < 3; ++d) m[r * 4 + d] = ax[r][d];
            m[r * 4 + 3] = -(ax[r][0] * c[0] + ax[r][1] * c[1] + ax[r][2] * c[2]);
        }
        m[12] = m[13] = m[14] = 0.f; m[15] = 1.f;
        float *dp = &depths[(size_t)k * H * W];
        for (int y = 0; y < H; ++y)
            for (int x = 0; x < W; ++x) dp[(size_t)y * W + x] = ((x / 60 + y / 60 + k) % 9 == 0) ? 0.f : 1.2f + 0.8f * sinf(0.004f * x + 0.3f * k) * cosf(0.006f * y);
    }
    // a fan of faces over consecutive vertices, and a box as the convex bound
    const int64_t T = V > 2 ? 2 * (V - 2) : 0;
    std::vector<int32_t> faces(T * 3);
    for (int64_t t = 0; t < T; ++t) { const int64_t b = t / 2; faces[t * 3] = (int32_t)b; faces[t * 3 + 1] = (int32_t)(b + 1 + (t & 1)); faces[t * 3 + 2] = (int32_t)((b * 7 + 3) % V); }
    const float planes_h[6][4] = {{1, 0, 0, -2.f}, {-1, 0, 0, -2.5f}, {0, 1, 0, -1.5f}, {0, -1, 0, -1.f}, {0, 0, 1, -1.f}, {0, 0, -1, -1.2f}};
    std::vector<float> planes(&planes_h[0][0], &planes_h[0][0] + 24);
    std::vector<uint8_t> colors(V * 3);
    for (int64_t i = 0; i < V * 3; ++i) colors[i] = (uint8_t)(i % 251);

    if (argc > 3 && !strcmp(argv[3], "host")) {        // no GPU: statistics of the synthetic problem from the host thread functions alone
        for (int eval_rec = 1; eval_rec >= 0; --eval_rec) {
            std::vector<uint8_t> s(V, 0);
            usl::CullFramesArgs A;
            A.verts = verts.data(); A.V = V; A.w2c = w2c.data(); A.depths = eval_rec ? depths.data() : nullptr; A.K = K; A.frames_per_cta = 16;
            A.cam = usl::CullCam{H, W, fx, fy, cx, cy, tr, eval_rec}; A.seen = s.data();
            for (int g = 0; g < (K + 15) / 16; ++g)
                for (int64_t tid = 0; tid < 4096; ++tid) usl::cull_frames_thread(A, tid, 4096, g);
            int64_t n = 0;
            for (int64_t v = 0; v < V; ++v) n += s[v];
            printf("host only: eval_rec %d seen %lld of %lld\n", eval_rec, (long long)n, (long long)V);
        }
        std::vector<uint8_t> in(V);
        for (int64_t tid = 0; tid < 4096; ++tid) usl::cull_hull_thread(verts.data(), V, planes.data(), 6, in.data(), tid, 4096);
        int64_t n = 0;
        for (int64_t v = 0; v < V; ++v) n += in[v];
        printf("host only: inside hull %lld of %lld\n", (long long)n, (long long)V);
        return 0;
    }
    float *d_verts = to_device(verts), *d_w2c = to_device(w2c), *d_depths = to_device(depths), *d_planes = to_device(planes);
    int32_t *d_faces = to_device(faces);
    uint8_t *d_colors = to_device(colors), *d_seen = nullptr, *d_inside = nullptr, *d_keep = nullptr, *d_vref = nullptr;
    if (!d_verts || !d_w2c || !d_depths || !d_planes || !d_faces || !d_colors) { printf("{\"error\": \"cudaMalloc\"}\n"); return 2; }
    CK(cudaMalloc(&d_seen, V + 16)); CK(cudaMalloc(&d_inside, V + 16)); CK(cudaMalloc(&d_keep, T + 16)); CK(cudaMalloc(&d_vref, V + 16));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    int bad = 0;
    printf("{\"vertices\": %lld, \"frames\": %d, \"faces\": %lld", (long long)V, K, (long long)T);

    // ---- usl_mesh_cull_frames, both eval_rec settings, vs the host thread functions ----
    std::vector<uint8_t> seen_h(V), seen_d(V), seen_rec(V);
    for (int eval_rec = 1; eval_rec >= 0; --eval_rec) {
        usl_cull_frames_args_t a;
        memset(&a, 0, sizeof(a));
        a.verts = d_verts; a.V = V; a.w2c = d_w2c; a.depths = d_depths; a.K = K; a.H = H; a.W = W;
        a.fx = fx; a.fy = fy; a.cx = cx; a.cy = cy; a.truncation = tr; a.eval_rec = eval_rec; a.frames_per_cta = 0; a.seen = d_seen;
        CK(cudaMemset(d_seen, 0, V)); ABI(usl_mesh_cull_frames(&a, nullptr)); CK(cudaDeviceSynchronize());       // warm-up
        CK(cudaMemset(d_seen, 0, V));
        CK(cudaEventRecord(e0)); ABI(usl_mesh_cull_frames(&a, nullptr)); CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
        float ms = 0; CK(cudaEventElapsedTime(&ms, e0, e1));
        CK(cudaMemcpy(seen_d.data(), d_seen, V, cudaMemcpyDeviceToHost));
        memset(seen_h.data(), 0, V);
        usl::CullFramesArgs A;
        A.verts = verts.data(); A.V = V; A.w2c = w2c.data(); A.depths = eval_rec ? depths.data() : nullptr; A.K = K; A.frames_per_cta = 16;
        A.cam = usl::CullCam{H, W, fx, fy, cx, cy, tr, eval_rec}; A.seen = seen_h.data();
        for (int g = 0; g < (K + 15) / 16; ++g)
            for (int64_t tid = 0; tid < 4096; ++tid) usl::cull_frames_thread(A, tid, 4096, g);
        int64_t mism = 0, n_seen = 0;
        for (int64_t v = 0; v < V; ++v) { mism += seen_h[v] != seen_d[v]; n_seen += seen_d[v]; }
        bad += mism != 0;
        printf(", \"cull_frames_%s\": {\"ms\": %.4f, \"seen\": %lld, \"mismatch_vs_host\": %lld}", eval_rec ? "occlusion" : "frustum_only", ms,
               (long long)n_seen, (long long)mism);
        if (eval_rec) seen_rec = seen_d;
    }

    // ---- usl_mesh_cull_hull ----
    std::vector<uint8_t> inside_h(V), inside_d(V);
    ABI(usl_mesh_cull_hull(d_verts, V, d_planes, 6, d_inside, nullptr)); CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0)); ABI(usl_mesh_cull_hull(d_verts, V, d_planes, 6, d_inside, nullptr)); CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
    float ms_h = 0; CK(cudaEventElapsedTime(&ms_h, e0, e1));
    CK(cudaMemcpy(inside_d.data(), d_inside, V, cudaMemcpyDeviceToHost));
    for (int64_t tid = 0; tid < 4096; ++tid) usl::cull_hull_thread(verts.data(), V, planes.data(), 6, inside_h.data(), tid, 4096);
    int64_t mism_h = 0, n_in = 0;
    for (int64_t v = 0; v < V; ++v) { mism_h += inside_h[v] != inside_d[v]; n_in += inside_d[v]; }
    bad += mism_h != 0;
    printf(", \"cull_hull\": {\"ms\": %.4f, \"inside\": %lld, \"mismatch_vs_host\": %lld}", ms_h, (long long)n_in, (long long)mism_h);

    // ---- face rule + scans + compaction (cull_mesh: keep a face when any vertex is seen) ----
    if (T > 0) {
        CK(cudaMemcpy(d_seen, seen_rec.data(), V, cudaMemcpyHostToDevice));
        int64_t nb = 0;
        ABI(usl_scan_u8_blocks(V > T ? V : T, &nb));
        uint32_t *d_sums = nullptr, *d_voff = nullptr, *d_foff = nullptr, *d_tot = nullptr;
        CK(cudaMalloc(&d_sums, (nb + 1) * 4)); CK(cudaMalloc(&d_voff, V * 4 + 16)); CK(cudaMalloc(&d_foff, T * 4 + 16)); CK(cudaMalloc(&d_tot, 8));
        CK(cudaEventRecord(e0));
        CK(cudaMemsetAsync(d_vref, 0, V, nullptr));
        ABI(usl_mesh_face_keep(d_faces, T, d_seen, V, 0, d_keep, d_vref, nullptr));
        ABI(usl_scan_u8(d_vref, V, 0, d_voff, d_sums, d_tot, nullptr));
        ABI(usl_scan_u8(d_keep, T, 0, d_foff, d_sums, d_tot + 1, nullptr));
        uint32_t tot[2] = {0, 0};
        CK(cudaMemcpy(tot, d_tot, 8, cudaMemcpyDeviceToHost));
        float *d_vo = nullptr; uint8_t *d_co = nullptr; int32_t *d_fo = nullptr;
        CK(cudaMalloc(&d_vo, (size_t)tot[0] * 12 + 16)); CK(cudaMalloc(&d_co, (size_t)tot[0] * 3 + 16)); CK(cudaMalloc(&d_fo, (size_t)tot[1] * 12 + 16));
        ABI(usl_mesh_compact(d_verts, d_colors, V, d_faces, T, d_keep, d_vref, d_voff, d_foff, d_vo, d_co, d_fo, nullptr));
        CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
        float ms_c = 0; CK(cudaEventElapsedTime(&ms_c, e0, e1));
        // host: the same thread functions + serial scans
        std::vector<uint8_t> keep_h(T), vref_h(V, 0);
        for (int64_t tid = 0; tid < 4096; ++tid) usl::cull_face_keep_thread(faces.data(), T, seen_rec.data(), V, 0, keep_h.data(), vref_h.data(), tid, 4096);
        std::vector<uint32_t> voff_h(V), foff_h(T);
        uint32_t nv = 0, nf = 0;
        for (int64_t v = 0; v < V; ++v) { voff_h[v] = nv; nv += vref_h[v]; }
        for (int64_t t = 0; t < T; ++t) { foff_h[t] = nf; nf += keep_h[t]; }
        std::vector<float> vo_h((size_t)nv * 3), vo_d((size_t)tot[0] * 3);
        std::vector<uint8_t> co_h((size_t)nv * 3), co_d((size_t)tot[0] * 3);
        std::vector<int32_t> fo_h((size_t)nf * 3), fo_d((size_t)tot[1] * 3);
        usl::CompactArgs C{verts.data(), colors.data(), V, faces.data(), T, keep_h.data(), vref_h.data(), voff_h.data(), foff_h.data(), vo_h.data(), co_h.data(), fo_h.data()};
        for (int64_t tid = 0; tid < 4096; ++tid) usl::cull_compact_thread(C, tid, 4096);
        int ok = (nv == tot[0]) && (nf == tot[1]);
        if (ok) {
            CK(cudaMemcpy(vo_d.data(), d_vo, vo_d.size() * 4, cudaMemcpyDeviceToHost));
            CK(cudaMemcpy(co_d.data(), d_co, co_d.size(), cudaMemcpyDeviceToHost));
            CK(cudaMemcpy(fo_d.data(), d_fo, fo_d.size() * 4, cudaMemcpyDeviceToHost));
            ok = !memcmp(vo_d.data(), vo_h.data(), vo_d.size() * 4) && !memcmp(co_d.data(), co_h.data(), co_d.size()) && !memcmp(fo_d.data(), fo_h.data(), fo_d.size() * 4);
        }
        bad += !ok;
        printf(", \"face_rule_scan_compact\": {\"ms\": %.4f, \"vertices_out\": %u, \"faces_out\": %u, \"equals_host\": %s}", ms_c, tot[0], tot[1], ok ? "true" : "false");
    }

    // ---- usl_render_metrics on one 1200x680 frame ----
    {
        const int64_t n = (int64_t)H * W;
        std::vector<float> gc(n * 3), c(n * 3), gd(depths.begin(), depths.begin() + n), d(n);
        for (int64_t i = 0; i < n * 3; ++i) { gc[i] = frand(); c[i] = gc[i] + 0.1f * (frand() - 0.5f); }
        for (int64_t i = 0; i < n; ++i) d[i] = gd[i] + 0.05f * (frand() - 0.5f);
        float *d_gc = to_device(gc), *d_c = to_device(c), *d_gd = to_device(gd), *d_d = to_device(d);
        double *d_acc = nullptr;
        CK(cudaMalloc(&d_acc, 24));
        CK(cudaMemset(d_acc, 0, 24)); ABI(usl_render_metrics(d_gc, d_gd, d_c, d_d, n, d_acc, nullptr)); CK(cudaDeviceSynchronize());
        CK(cudaMemset(d_acc, 0, 24));
        CK(cudaEventRecord(e0)); ABI(usl_render_metrics(d_gc, d_gd, d_c, d_d, n, d_acc, nullptr)); CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
        float ms_m = 0; CK(cudaEventElapsedTime(&ms_m, e0, e1));
        double acc_d[3], acc_h[3] = {0, 0, 0};
        CK(cudaMemcpy(acc_d, d_acc, 24, cudaMemcpyDeviceToHost));
        for (int64_t tid = 0; tid < 4096; ++tid) {
            double se, ad, cnt;
            usl::metrics_thread(gc.data(), gd.data(), c.data(), d.data(), n, tid, 4096, se, ad, cnt);
            acc_h[0] += se; acc_h[1] += ad; acc_h[2] += cnt;
        }
        const double r0 = fabs(acc_d[0] - acc_h[0]) / acc_h[0], r1 = fabs(acc_d[1] - acc_h[1]) / acc_h[1];
        const int ok = acc_d[2] == acc_h[2] && r0 < 1e-12 && r1 < 1e-12;
        bad += !ok;
        printf(", \"render_metrics\": {\"ms\": %.4f, \"gb_per_s\": %.1f, \"pixels_with_depth\": %.0f, \"rel_diff_vs_host\": %.3e, \"ok\": %s}", ms_m,
               n * 32 / (ms_m * 1e-3) / 1e9, acc_d[2], r0 > r1 ? r0 : r1, ok ? "true" : "false");
    }
    printf(", \"all_equal_host\": %s}\n", bad ? "false" : "true");
    return bad ? 1 : 0;
}
