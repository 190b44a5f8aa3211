"""CPU: the committed fixtures come out of the committed generator.  Re-runs oracle/gen_golden.py (the UNMODIFIED
reference through oracle/shims) for two cases into a temporary directory and compares every array bit for bit with
tests/golden/.  Skipped where the reference tree is absent (the GPU box)."""
import os
import subprocess
import sys

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASES = ["map_scannet_k23", "map_replica_k7", "track_scannet", "mesh_replica", "cull_replica", "evalr_replica"]


@pytest.mark.skipif(not os.path.isdir("/root/reference/src"), reason="needs the reference tree (build container only)")
def test_generator_reproduces_committed_fixtures(tmp_path):
    env = dict(os.environ, USL_GOLDEN_OUT=str(tmp_path))
    subprocess.run([sys.executable, "-m", "oracle.gen_golden"] + CASES, cwd=REPO, env=env, check=True,
                   stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, timeout=600)
    for name in CASES:
        new = np.load(os.path.join(str(tmp_path), name + ".npz"), allow_pickle=False)
        old = np.load(os.path.join(REPO, "tests", "golden", name + ".npz"), allow_pickle=False)
        assert sorted(new.files) == sorted(old.files), name
        for k in old.files:
            a, b = old[k], new[k]
            assert a.dtype == b.dtype and a.shape == b.shape, (name, k)
            assert a.tobytes() == b.tobytes(), (name, k)          # bit-equal (NaN-safe)


def test_generator_does_not_import_the_product_package():
    """Test infrastructure must not depend on product code: the generator's scene lives in oracle/scene.py."""
    for f in ("gen_golden.py", "scene.py", "time_reference_cpu.py"):
        src = open(os.path.join(REPO, "oracle", f)).read()
        assert "import_module" not in src and "from uni" not in src, f
