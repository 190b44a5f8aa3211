"""ctypes binding of lib/libunislam_b200.so (the C-ABI declared in include/unislam_b200.h).

There is NO fallback: if the shared library is missing or a call fails, a RuntimeError is raised
(tinycudann behaves the same way when its extension is absent).  Pointers are raw
``tensor.data_ptr()`` values; the stream is ``torch.cuda.current_stream().cuda_stream`` so every
kernel runs on torch's current stream and is CUDA-graph capturable.
"""
import ctypes
import os
from ctypes import POINTER, Structure, byref, c_char_p, c_double, c_float, c_int, c_int32, c_int64, c_uint8, c_uint32, c_void_p

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# USL_LIB_PATH: development only -- load a side-by-side build variant (csrc/build.sh with USL_LIB_NAME) for A/B measurements
LIB_PATH = os.environ.get("USL_LIB_PATH") or os.path.join(_HERE, "lib", "libunislam_b200.so")

MAX_LEVELS = 16
ACT_NONE, ACT_TANH, ACT_SIGMOID = 0, 1, 2
LOSS_SLOTS = 16


class Level(Structure):
    _fields_ = [("scale", c_float), ("res", c_uint32), ("size", c_uint32), ("offset", c_uint32), ("hashed", c_uint32)]


class Grid(Structure):
    _fields_ = [("n_levels", c_int32), ("total_entries", c_uint32), ("levels", Level * MAX_LEVELS)]


class Mlp(Structure):
    _fields_ = [("w1", c_void_p), ("b1", c_void_p), ("w2", c_void_p), ("b2", c_void_p), ("wo", c_void_p), ("bo", c_void_p),
                ("n_hidden", c_int32), ("n_out", c_int32), ("out_act", c_int32), ("_pad", c_int32)]


class Field(Structure):
    _fields_ = [("grid", Grid * 2), ("table", c_void_p * 2), ("mlp", Mlp * 2), ("bound_lo", c_float * 3), ("bound_hi", c_float * 3)]


class Bound(Structure):
    _fields_ = [("lo", c_float * 3), ("hi", c_float * 3)]


class Points(Structure):
    _fields_ = [("x", c_void_p), ("rays_o", c_void_p), ("rays_d", c_void_p), ("z", c_void_p), ("valid", c_void_p),
                ("S", c_int32), ("sample_major", c_int32), ("n", c_int64)]


class ZSampleArgs(Structure):
    _fields_ = [("n_stratified", c_int32), ("n_importance", c_int32), ("c_surf_lo", c_float), ("c_surf_span", c_float),
                ("t_uni", c_void_p), ("t_surf", c_void_p)]


class RayBatch(Structure):
    _fields_ = [("c2ws", c_void_p), ("depths", c_void_p), ("colors", c_void_p), ("dirs_cam", c_void_p), ("indices", c_void_p),
                ("P", c_int64), ("K", c_int32), ("n", c_int32), ("frame_base", c_int32), ("_pad", c_int32)]


class RaySetup(Structure):
    _fields_ = [("mode", c_int32), ("n_batches", c_int32), ("batch", RayBatch * 2),
                ("depth_img", c_void_p), ("color_img", c_void_p), ("win_indices", c_void_p),
                ("H", c_int32), ("W", c_int32), ("H0", c_int32), ("H1", c_int32), ("W0", c_int32), ("W1", c_int32),
                ("fx", c_float), ("fy", c_float), ("cx", c_float), ("cy", c_float),
                ("c2w", c_void_p), ("cam_poses", c_void_p), ("c2w_fixed", c_void_p),
                ("bound", Bound), ("require_depth", c_int32), ("_pad2", c_int32),
                ("zs", ZSampleArgs), ("t_rand", c_void_p), ("n_rays", c_int64),
                ("rays_o", c_void_p), ("rays_d", c_void_p), ("gt_depth", c_void_p), ("gt_color", c_void_p), ("dirs_out", c_void_p),
                ("frame_id", c_void_p), ("valid", c_void_p), ("z", c_void_p), ("pixel_begin", c_int64), ("ray_offset", c_int64)]


class LossArgs(Structure):
    _fields_ = [("truncation", c_float), ("truncation_center", c_float), ("w_sdf_fs", c_float), ("w_sdf_center", c_float),
                ("w_sdf_tail", c_float), ("w_depth", c_float), ("w_color", c_float), ("mode", c_int32)]


class AdamGroup(Structure):
    _fields_ = [("param", c_void_p), ("grad", c_void_p), ("exp_avg", c_void_p), ("exp_avg_sq", c_void_p), ("n", c_int64),
                ("lr", c_float), ("beta1", c_float), ("beta2", c_float), ("eps", c_float), ("step", c_int64)]


class Peers(Structure):
    _fields_ = [("rank", c_int32), ("world", c_int32), ("buf", c_void_p * 8), ("ctrl", c_void_p * 8), ("mc", c_void_p),
                ("channel", c_int32), ("max_ctas_per_sm", c_int32)]


class AdamRange(Structure):
    _fields_ = [("begin", c_int64), ("end", c_int64), ("lr", c_float), ("_pad", c_float)]


class McArgs(Structure):
    _fields_ = [("vol", c_void_p), ("nx", c_int32), ("nz", c_int32), ("rows", c_int32), ("own_rows", c_int32), ("y_begin", c_int32),
                ("level", c_float), ("origin", c_float * 3), ("spacing", c_float * 3), ("pflags", c_void_p), ("ctri", c_void_p),
                ("voff", c_void_p), ("toff", c_void_p), ("verts", c_void_p), ("vkeys", c_void_p), ("faces", c_void_p)]


class CullFramesArgs(Structure):
    _fields_ = [("verts", c_void_p), ("V", c_int64), ("w2c", c_void_p), ("depths", c_void_p), ("K", c_int32), ("H", c_int32), ("W", c_int32),
                ("fx", c_float), ("fy", c_float), ("cx", c_float), ("cy", c_float), ("truncation", c_float), ("eval_rec", c_int32),
                ("frames_per_cta", c_int32), ("seen", c_void_p)]


ADAM_MAX_GROUPS = 24
_P = c_void_p
_SIGS = {
    "usl_grid_build": [c_int, c_int, c_int, c_double, POINTER(Grid)],
    "usl_grid_encode_fwd": [POINTER(Grid), _P, _P, c_int64, _P, _P],
    "usl_grid_encode_bwd_params": [POINTER(Grid), _P, _P, c_int64, _P, _P],
    "usl_grid_encode_bwd_input": [POINTER(Grid), _P, _P, _P, c_int64, _P, _P],
    "usl_grid_corner_indices": [POINTER(Grid), _P, c_int64, _P, _P],
    "usl_mlp_fwd": [POINTER(Mlp), _P, c_int64, _P, _P],
    "usl_mlp_bwd": [POINTER(Mlp), POINTER(Mlp), _P, _P, _P, c_int64, _P, _P],
    "usl_sample_keyframe_rays": [_P, _P, _P, _P, _P, c_int, c_int64, c_int, c_int, _P, _P, _P, _P, _P, _P, _P],
    "usl_sample_window_rays": [_P, _P, _P, c_int, c_int, c_int, c_int, c_int, c_int, c_float, c_float, c_float, c_float,
                               _P, c_int64, _P, _P, _P, _P, _P, _P],
    "usl_image_rays": [_P, c_int, c_int, c_float, c_float, c_float, c_float, _P, _P, _P],
    "usl_bbox_prefilter": [_P, _P, _P, c_int64, POINTER(Bound), c_int, _P, _P, _P],
    "usl_zsample_depth": [POINTER(ZSampleArgs), _P, _P, _P, _P, c_int64, _P, _P],
    "usl_zsample_nodepth": [POINTER(ZSampleArgs), POINTER(Field), _P, _P, _P, _P, _P, _P, _P, _P, c_int64, _P, _P, _P],
    "usl_field_fwd": [POINTER(Field), POINTER(Points), _P, _P, _P, _P],
    "usl_field_fwd_tc": [POINTER(Field), POINTER(Points), _P, _P, _P, _P],
    "usl_field_bwd": [POINTER(Field), POINTER(Points), _P, _P, _P, _P, _P, POINTER(Mlp), _P, c_int, _P],
    "usl_field_bwd_scratch_floats": [POINTER(Field), POINTER(c_int64)],
    "usl_field_stash_floats": [c_int64, POINTER(c_int64)],
    "usl_field_sdf": [POINTER(Field), POINTER(Points), _P, _P],
    "usl_composite_fwd": [_P, _P, _P, _P, c_int64, c_int, _P, _P, _P, _P, _P, _P, _P],
    "usl_composite_bwd": [_P, _P, _P, _P, c_int64, c_int, _P, _P, _P, _P, _P, _P, _P, POINTER(Bound), _P, _P, _P, _P, _P],
    "usl_composite_loss_fwd": [POINTER(LossArgs), _P, _P, _P, _P, c_int64, c_int, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P],
    "usl_loss_fwd": [POINTER(LossArgs), _P, _P, _P, _P, _P, _P, _P, _P, _P, c_int64, c_int, _P, _P, _P],
    "usl_loss_finalize": [POINTER(LossArgs), _P, _P, _P],
    "usl_loss_bwd": [POINTER(LossArgs), _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, c_int64, c_int, _P, _P, _P, _P],
    "usl_depth_error_median": [_P, _P, _P, c_int64, _P, _P, _P],
    "usl_pose_reduce": [_P, _P, _P, _P, _P, c_int64, c_int, _P, _P],
    "usl_pose_to_matrix": [_P, c_int, _P, _P],
    "usl_pose_matrix_bwd": [_P, _P, c_int, _P, _P],
    "usl_track_keep_best": [_P, _P, _P, _P, _P],
    "usl_sdf_query_grid": [POINTER(Field), _P, _P, _P, c_int, c_int, c_int, c_int, c_int, _P, _P],
    "usl_ray_setup": [POINTER(RaySetup), _P],
    "usl_composite_loss_bwd": [POINTER(LossArgs), _P, _P, _P, _P, _P, c_int64, c_int, _P, _P, _P, _P, _P, _P, _P, POINTER(Bound),
                               _P, _P, _P, _P, _P, _P],
    "usl_adam_step": [POINTER(AdamGroup), c_int, c_int64, _P, c_int, _P],
    "usl_keyframe_insert": [_P, _P, _P, _P, c_int64, _P, _P, _P, _P, _P],
    "usl_keyframe_covisibility": [_P, _P, _P, c_int64, c_int, _P, c_int, c_int, c_int, c_float, c_float, c_float, c_float, c_float, _P, _P],
    "usl_mc_classify": [POINTER(McArgs), _P],
    "usl_mc_emit": [POINTER(McArgs), _P],
    "usl_scan_u8": [_P, c_int64, c_int, _P, _P, _P, _P],
    "usl_scan_u8_blocks": [c_int64, POINTER(c_int64)],
    "usl_render_metrics": [_P, _P, _P, _P, c_int64, _P, _P],
    "usl_mesh_cull_frames": [POINTER(CullFramesArgs), _P],
    "usl_mesh_cull_hull": [_P, c_int64, _P, c_int32, _P, _P],
    "usl_mesh_face_keep": [_P, c_int64, _P, c_int64, c_int32, _P, _P, _P],
    "usl_mesh_compact": [_P, _P, c_int64, _P, c_int64, _P, _P, _P, _P, _P, _P, _P, _P],
    "usl_exchange_sums": [POINTER(Peers), _P, _P],
    "usl_peer_barrier": [POINTER(Peers), _P],
    "usl_allreduce_sum": [POINTER(Peers), c_int64, c_int64, _P],
    "usl_allreduce_adam_slice_floats": [c_int, c_int64, POINTER(c_int64)],
    "usl_allreduce_adam_step": [POINTER(Peers), POINTER(c_void_p), _P, c_int64, c_int64, _P, _P, POINTER(AdamRange), c_int, c_float, c_float, c_float,
                                c_int64, _P, _P],
    "usl_bench_stream_read": [_P, c_int64, c_int, _P, _P],
    "usl_bench_gather": [_P, c_uint32, c_int64, c_int, _P, _P],
    "usl_bench_scatter": [_P, c_uint32, c_int64, c_int, c_int, _P],
}
EXPORTS = ["usl_last_error", "usl_version", "usl_peer_ctrl_bytes"] + list(_SIGS)

_lib = None


def load():
    """Load the C-ABI library; raise loudly when it is absent (no CPU / eager fallback exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"unislam_b200: CUDA library not built: {LIB_PATH} is missing. Run `python -c \"import __graft_entry__ as g; "
            f"g.build()\"` (or uni-slam_b200/csrc/build.sh). There is no CPU fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    lib.usl_last_error.restype = c_char_p
    lib.usl_last_error.argtypes = []
    lib.usl_version.restype = c_int
    lib.usl_version.argtypes = []
    lib.usl_peer_ctrl_bytes.restype = c_int
    lib.usl_peer_ctrl_bytes.argtypes = []
    for name, sig in _SIGS.items():
        fn = getattr(lib, name)
        fn.restype = c_int
        fn.argtypes = sig
    _lib = lib
    return lib


LAUNCHES = 0   # number of kernel-launching C-ABI calls made so far (bench.py reports it as gpu_launches)


def call(name, *args):
    global LAUNCHES
    lib = load()
    if name not in ("usl_grid_build", "usl_field_bwd_scratch_floats", "usl_field_stash_floats", "usl_allreduce_adam_slice_floats", "usl_scan_u8_blocks"):
        LAUNCHES += 1
    rc = getattr(lib, name)(*args)
    if rc != 0:
        raise RuntimeError(f"{name} failed: {lib.usl_last_error().decode()}")


def stream():
    return torch.cuda.current_stream().cuda_stream


def ptr(t):
    """Device pointer of a contiguous CUDA tensor (None -> NULL)."""
    if t is None:
        return None
    assert t.is_cuda and t.is_contiguous(), "unislam_b200 ops need contiguous CUDA tensors"
    return t.data_ptr()


def cptr(t, dtype, numel=None, name="tensor"):
    """Checked device pointer for CALLER-supplied tensors: the kernels reinterpret raw memory, so a wrong dtype, a CPU
    tensor, a strided view or a short buffer must be refused here (None -> NULL)."""
    if t is None:
        return None
    if not (t.is_cuda and t.is_contiguous()):
        raise ValueError(f"{name}: need a contiguous CUDA tensor")
    if t.dtype != dtype:
        raise ValueError(f"{name}: dtype {t.dtype}, expected {dtype}")
    if numel is not None and t.numel() < numel:
        raise ValueError(f"{name}: {t.numel()} elements, need at least {numel}")
    return t.data_ptr()


def f32c(t):
    """fp32, contiguous (the tcnn torch binding does the same cast)."""
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def build_grid(n_levels, log2_hashmap_size, base_resolution, per_level_scale) -> Grid:
    g = Grid()
    call("usl_grid_build", int(n_levels), int(log2_hashmap_size), int(base_resolution), float(per_level_scale), byref(g))
    return g


def make_bound(bound) -> Bound:
    b = Bound()
    bb = bound.detach().to("cpu", torch.float32)
    for d in range(3):
        b.lo[d] = float(bb[d, 0]); b.hi[d] = float(bb[d, 1])
    return b
