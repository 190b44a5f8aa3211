"""oracle/ -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU restatement of the Uni-SLAM differentiable-rendering hot path:

* ``grid_oracle.c``  plain-C restatement of tiny-cuda-nn's HashGrid encoding
                     (PARITY UNPINNED for the tcnn arithmetic: its source is an un-vendored
                     dependency, requirements.txt:90, absent from /root/reference).
* ``grid_ref.py``    the same encoding as differentiable fp32/fp64 PyTorch ops.
* ``path_ref.py``    restatement of src/utils/Renderer.py, src/common.py,
                     src/networks/decoders.py and the loss code of src/Mapper.py /
                     src/Tracker.py.  Pinned against the reference's own Python code run in
                     this container through ``oracle/shims`` (see ``gen_golden.py`` and
                     ``tests/golden/``).
* ``shims/``         import stubs that let the UNMODIFIED reference run on CPU here
                     (tinycudann -> grid_ref, pytorch3d.transforms -> restated quaternion math).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import anything from this package.  The product package
(``uni-slam_b200``) never does and fails loudly when its CUDA library is missing.
"""
import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_BUILD = os.path.join(_HERE, "_build")
_LIB = os.path.join(_BUILD, "liboracle_grid.so")
_lib = None


def build(force: bool = False) -> str:
    """Compile grid_oracle.c with gcc (a few 100 ms). Output: oracle/_build/liboracle_grid.so."""
    src = os.path.join(_HERE, "grid_oracle.c")
    if (not force) and os.path.exists(_LIB) and os.path.getmtime(_LIB) >= os.path.getmtime(src):
        return _LIB
    os.makedirs(_BUILD, exist_ok=True)
    # -ffp-contract=off: every fmaf in the source is explicit; nothing else may be fused.
    cmd = ["gcc", "-O2", "-std=c11", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math",
           "-o", _LIB, src, "-lm"]
    subprocess.run(cmd, check=True)
    return _LIB


class OrcLevel(ctypes.Structure):
    _fields_ = [("scale", ctypes.c_float), ("res", ctypes.c_uint32), ("size", ctypes.c_uint32),
                ("offset", ctypes.c_uint32), ("hashed", ctypes.c_uint32)]


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB)
        L.orc_grid_levels.restype = ctypes.c_int
        L.orc_grid_levels.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_double,
                                      ctypes.POINTER(OrcLevel), ctypes.POINTER(ctypes.c_uint32)]
        L.orc_grid_index.restype = ctypes.c_uint32
        L.orc_grid_index.argtypes = [ctypes.POINTER(OrcLevel), ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32]
        vp, i64 = ctypes.c_void_p, ctypes.c_int64
        L.orc_grid_corners.restype = None
        L.orc_grid_corners.argtypes = [ctypes.POINTER(OrcLevel), ctypes.c_int, vp, i64, vp, vp]
        L.orc_grid_encode_fwd.restype = None
        L.orc_grid_encode_fwd.argtypes = [ctypes.POINTER(OrcLevel), ctypes.c_int, vp, vp, i64, vp]
        L.orc_grid_encode_bwd_params.restype = None
        L.orc_grid_encode_bwd_params.argtypes = [ctypes.POINTER(OrcLevel), ctypes.c_int, vp, vp, i64, vp]
        L.orc_grid_encode_bwd_input.restype = None
        L.orc_grid_encode_bwd_input.argtypes = [ctypes.POINTER(OrcLevel), ctypes.c_int, vp, vp, vp, i64, vp]
        _lib = L
    return _lib
