"""Wiring of this package under the UNMODIFIED reference tree.

``install()`` puts ``compat/`` (a ``tinycudann`` module backed by the C-ABI library) ahead of any real
tiny-cuda-nn on ``sys.path`` and, optionally, swaps the reference's ``Decoders`` / ``Renderer`` classes for
the fused drop-ins before ``src.UNISLAM`` binds them.  ``src/Mapper.py``, ``src/Tracker.py`` and
``src/common.py`` stay byte-identical."""
import importlib
import importlib.util
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))


def install(reference_root=None, fused=True):
    """reference_root: path of the Uni-SLAM checkout (added to sys.path when given).
    fused=False keeps the reference's own Decoders / Renderer and only replaces tinycudann (seams B1/B2);
    fused=True also routes Decoders.forward / Renderer.render_batch_ray through the fused kernels (B3/B4)."""
    compat = os.path.join(_HERE, "compat")
    if compat not in sys.path:
        sys.path.insert(0, compat)
    if importlib.util.find_spec("pytorch3d") is None:
        p3d = os.path.join(_HERE, "compat_pytorch3d")
        if p3d not in sys.path:
            sys.path.insert(1, p3d)
    if reference_root and reference_root not in sys.path:
        sys.path.insert(2, reference_root)
    if fused:
        modules = importlib.import_module(__package__ + ".modules")
        dec_mod = importlib.import_module("src.networks.decoders")
        ren_mod = importlib.import_module("src.utils.Renderer")
        dec_mod.Decoders = modules.Decoders          # src/networks/config.py:18 imports it from here
        ren_mod.Renderer = modules.Renderer          # src/UNISLAM.py:32 imports it from here
        cfg_mod = sys.modules.get("src.networks.config")
        if cfg_mod is not None:
            cfg_mod.Decoders = modules.Decoders
        uni = sys.modules.get("src.UNISLAM")
        if uni is not None:
            uni.Renderer = modules.Renderer
