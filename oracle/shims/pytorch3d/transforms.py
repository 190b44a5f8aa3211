"""pytorch3d.transforms shim -- TEST INFRASTRUCTURE (pure-PyTorch functions restated)."""
from oracle.path_ref import matrix_to_quaternion, quaternion_to_matrix  # noqa: F401
