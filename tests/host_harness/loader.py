"""TEST INFRASTRUCTURE, NOT PRODUCT CODE: loader of tests/host_harness/cull_host.cpp.

The harness compiles uni-slam_b200/csrc/usl_cull.cuh -- the thread functions the CUDA kernels of cull.cu inline -- for the
host with g++ (-ffp-contract=off), so that their arithmetic can be checked without a GPU and the kernels can then be held to
it bit for bit.  Imported by tests/ (through helpers.py) and by bench.py's isolated culling leg as the checker; it imports
nothing from oracle/ and nothing from the product package, and the product never loads it.
"""
import os

import numpy as np

_CULL_HOST = None
_METRICS_HOST = None


def _build(src_name, hdr_name, out_name):
    """g++ build into tests/_build/ when the library is missing or older than its sources (to a temporary name, then renamed:
    concurrent test processes never see a half-written file).  If the compiler fails but a library from an earlier build
    exists, that one is used and the mismatch, if any, shows up in the comparison it serves."""
    import subprocess
    import warnings
    here = os.path.dirname(os.path.abspath(__file__))
    tests = os.path.dirname(here)
    src = os.path.join(here, src_name)
    hdr = os.path.join(os.path.dirname(tests), "uni-slam_b200", "csrc", hdr_name)
    out = os.path.join(tests, "_build", out_name)
    if not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        os.makedirs(os.path.dirname(out), exist_ok=True)
        tmp = f"{out}.{os.getpid()}.tmp"
        try:
            subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math", "-o", tmp, src], check=True)
            os.replace(tmp, out)
        except (OSError, subprocess.CalledProcessError) as e:
            if not os.path.exists(out):
                raise
            warnings.warn(f"host harness: rebuild of {out_name} failed ({e}); using the existing library")
    return out


def gpu_check_binary():
    """tests/_build/cull_gpu_check: the torch-free C-ABI consumer of cull_gpu_check.cu (nvcc, linked against the product library
    with an rpath relative to the binary).  Returns its path."""
    import subprocess
    here = os.path.dirname(os.path.abspath(__file__))
    tests = os.path.dirname(here)
    repo = os.path.dirname(tests)
    src = os.path.join(here, "cull_gpu_check.cu")
    deps = [src] + [os.path.join(repo, "uni-slam_b200", "csrc", h) for h in ("usl_cull.cuh", "usl_metrics.cuh")] + [os.path.join(repo, "include", "unislam_b200.h")]
    out = os.path.join(tests, "_build", "cull_gpu_check")
    if not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(d) for d in deps):
        os.makedirs(os.path.dirname(out), exist_ok=True)
        tmp = f"{out}.{os.getpid()}.tmp"
        try:
            subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O2", "-std=c++17", "-cudart", "shared", "-Xcompiler", "-ffp-contract=off", "-o", tmp, src,
                            "-I" + os.path.join(repo, "include"), "-L" + os.path.join(repo, "uni-slam_b200", "lib"), "-lunislam_b200",
                            "-Xlinker", "-rpath", "-Xlinker", "$ORIGIN/../../uni-slam_b200/lib"], check=True)
            os.replace(tmp, out)
        except (OSError, subprocess.CalledProcessError):
            if not os.path.exists(out):
                raise
    return out


def metrics_host(gt_color, gt_depth, color, depth, nthreads=148 * 4 * 256):
    """acc (3,) float64 = [sum of squared colour errors, sum |depth error|, pixel count] over the pixels with gt_depth > 0, from
    the host build of usl_metrics.cuh (tests/host_harness/metrics_host.cpp)."""
    global _METRICS_HOST
    import ctypes
    if _METRICS_HOST is None:
        lib = ctypes.CDLL(_build("metrics_host.cpp", "usl_metrics.cuh", "libmetrics_host.so"))
        lib.metrics_host.restype = None
        lib.metrics_host.argtypes = [ctypes.c_void_p] * 4 + [ctypes.c_int64, ctypes.c_int64, ctypes.c_void_p]
        _METRICS_HOST = lib
    a = [np.ascontiguousarray(x, dtype=np.float32) for x in (gt_color, gt_depth, color, depth)]
    acc = np.zeros(3, dtype=np.float64)
    _METRICS_HOST.metrics_host(a[0].ctypes.data, a[1].ctypes.data, a[2].ctypes.data, a[3].ctypes.data, a[1].size, int(nthreads), acc.ctypes.data)
    return acc


def cull_host():
    """g++ build of tests/host_harness/cull_host.cpp (it includes uni-slam_b200/csrc/usl_cull.cuh, the source the CUDA
    kernels inline) -> tests/_build/libcull_host.so.  Test infrastructure: never loaded by the product."""
    global _CULL_HOST
    if _CULL_HOST is not None:
        return _CULL_HOST
    import ctypes
    lib = ctypes.CDLL(_build("cull_host.cpp", "usl_cull.cuh", "libcull_host.so"))
    vp, i64, ci, cf = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_float
    lib.cull_host_frames.restype = None
    lib.cull_host_frames.argtypes = [vp, i64, vp, vp, ci, ci, ci, cf, cf, cf, cf, cf, ci, ci, i64, vp]
    lib.cull_host_hull.restype = None
    lib.cull_host_hull.argtypes = [vp, i64, vp, ci, i64, vp]
    lib.cull_host_compact.restype = None
    lib.cull_host_compact.argtypes = [vp, vp, i64, vp, i64, vp, ci, i64, vp, vp, vp, vp, vp]
    _CULL_HOST = lib
    return lib


NTHREADS = 4736 * 256         # the x extent usl_mesh_cull_frames launches on a 148-SM device when there are more tiles than that


def cull_host_frames(verts, w2c, depths, cam, truncation, eval_rec, frames_per_cta=16, nthreads=NTHREADS):
    """seen (V,) uint8 from the host harness; verts (V,3) f32, w2c (K,4,4) f32, depths (K,H,W) f32, cam = (H,W,fx,fy,cx,cy);
    nthreads = simulated gridDim.x * blockDim.x."""
    lib = cull_host()
    verts = np.ascontiguousarray(verts, dtype=np.float32); w2c = np.ascontiguousarray(w2c, dtype=np.float32)
    depths = np.ascontiguousarray(depths, dtype=np.float32)
    seen = np.zeros(len(verts), dtype=np.uint8)
    H, W, fx, fy, cx, cy = cam
    lib.cull_host_frames(verts.ctypes.data, len(verts), w2c.ctypes.data, depths.ctypes.data, len(w2c), int(H), int(W), fx, fy, cx, cy,
                         truncation, int(eval_rec), frames_per_cta, int(nthreads), seen.ctypes.data)
    return seen


def cull_host_hull(verts, planes, nthreads=NTHREADS):
    lib = cull_host()
    verts = np.ascontiguousarray(verts, dtype=np.float32); planes = np.ascontiguousarray(planes, dtype=np.float32)
    inside = np.zeros(len(verts), dtype=np.uint8)
    lib.cull_host_hull(verts.ctypes.data, len(verts), planes.ctypes.data, len(planes), int(nthreads), inside.ctypes.data)
    return inside


def cull_host_compact(verts, colors, faces, vmask, require_all, nthreads=NTHREADS):
    lib = cull_host()
    verts = np.ascontiguousarray(verts, dtype=np.float32); faces = np.ascontiguousarray(faces, dtype=np.int32)
    vmask = np.ascontiguousarray(vmask, dtype=np.uint8)
    colors = np.ascontiguousarray(colors, dtype=np.uint8) if colors is not None else None
    keep = np.zeros(len(faces), dtype=np.uint8)
    vo = np.zeros_like(verts); fo = np.zeros_like(faces); co = np.zeros_like(colors) if colors is not None else None
    n = np.zeros(2, dtype=np.int64)
    lib.cull_host_compact(verts.ctypes.data, colors.ctypes.data if colors is not None else None, len(verts), faces.ctypes.data, len(faces),
                          vmask.ctypes.data, int(require_all), int(nthreads), keep.ctypes.data, vo.ctypes.data, co.ctypes.data if co is not None else None,
                          fo.ctypes.data, n.ctypes.data)
    return vo[:n[0]], fo[:n[1]], (co[:n[0]] if co is not None else None), keep
