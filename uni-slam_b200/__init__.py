"""unislam_b200 -- B200 (sm_100a) replacement for Uni-SLAM's per-frame differentiable-rendering hot path.

Import name note: the package directory is ``uni-slam_b200`` (hyphenated, as the build spec names it);
``import unislam_b200`` (the alias module at the repo root) or
``importlib.import_module("uni-slam_b200")`` both give this package.

Public surface (mirrors the reference's operator API for the path):
  Encoding, Network         tinycudann-compatible modules          (B1, B2)
  Decoders, Renderer        src/networks/decoders.py, src/utils/Renderer.py drop-ins (B3, B4)
  MappingStep, TrackingStep fused per-iteration drivers            (B5, B6 + a-3 .. a-10)
  DenseSdfQuery             dense SDF query for meshing            (a-11)
  RenderImageStep           forward-only whole-frame renderer      (f3, Renderer.render_img)
  RenderMetrics             eval_rendering's per-frame PSNR / depth L1 on the device (f3, tools/eval_recon.py:276-299)
  FusedAdam                 one-launch torch.optim.Adam equivalent (a-12 / f1)
  KeyframeStore             device-resident keyframe subsets + window views (f2, Mapper.py:315-356,528-541)
  mesh                      marching cubes on the device-resident volume, vertex colours, frame / bound culling, PLY
                            (f4, Mesher.py:230-276, tools/cull_mesh.py)
  parallel                  multi-GPU: peer-memory exchange kernels, slab / ray-range sharding (8e)
  ops                       thin per-kernel wrappers over the C-ABI
There is no CPU path: every op raises RuntimeError if lib/libunislam_b200.so is missing.
"""
from . import _lib, mesh, ops, synthetic  # noqa: F401
from .modules import Decoders, Encoding, Network, Renderer  # noqa: F401
from .keyframes import KeyframeStore  # noqa: F401
from .optim import FusedAdam  # noqa: F401
from .steps import DenseSdfQuery, MappingStep, RenderImageStep, RenderMetrics, TrackingStep  # noqa: F401

__all__ = ["Encoding", "Network", "Decoders", "Renderer", "MappingStep", "TrackingStep", "DenseSdfQuery", "RenderImageStep", "RenderMetrics", "FusedAdam", "KeyframeStore", "ops", "mesh", "synthetic"]
