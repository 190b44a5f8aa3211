"""``import tinycudann as tcnn`` resolves here when uni-slam_b200/compat is put on sys.path ahead of a
real tiny-cuda-nn install: the reference's src/UNISLAM.py:25,242-253 and
src/networks/decoders.py:22,50-70 then construct the B200 modules without any source change."""
import importlib

_m = importlib.import_module("uni-slam_b200.modules")
Encoding = _m.Encoding
Network = _m.Network
__all__ = ["Encoding", "Network"]
