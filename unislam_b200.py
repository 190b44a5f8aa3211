"""Import alias: ``import unislam_b200`` -> the package in ./uni-slam_b200/ (hyphenated directory name)."""
import importlib
import sys

_pkg = importlib.import_module("uni-slam_b200")
sys.modules[__name__] = _pkg
