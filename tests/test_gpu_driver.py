"""The C++ driver shim (stand-in for the patched mcpolar.f90) against libtamc.so: BASELINE config 5's
call pattern -- repeated set_optics + run with a growing crater -- and writer.f90's output format."""
import os
import re
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "tissue-ablation-mc_b200")


def test_coupled_loop_shim(tmp_path):
    exe = os.path.join(PKG, "mcgrid_shim")
    if not os.path.exists(exe):
        pytest.skip("driver shim not built")
    res = subprocess.run([exe, "--params", os.path.join(PKG, "driver", "input.params.example"), "--calls", "12",
                          "--out", str(tmp_path)], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stderr
    assert "# of photons to run 125000 per call" in res.stdout
    m = re.search(r"packets (\d+), voxel-steps (\d+)", res.stdout)
    assert m and int(m.group(1)) == 12 * 125000
    # the crater deepens: more voxels crossed per packet than the intact 1.564
    assert int(m.group(2)) / int(m.group(1)) > 1.6
    f = tmp_path / "jmean-t70w-80-500-400-0.030-0.030-0.060.dat"      # writer.f90:23-25 naming
    assert f.exists() and f.stat().st_size == 80 ** 3 * 8
    jm = np.fromfile(f, dtype="<f8").reshape((80, 80, 80), order="F")
    assert jm.min() >= 0 and jm.sum() > 0
    assert jm[39:41, 39:41, 79].sum() == 0.0                            # ablated centre takes no deposit
