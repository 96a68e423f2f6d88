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


def test_resident_loop_and_writer_files(tmp_path):
    """--resident: the reference's time loop with its real heat step on the device, then writer.f90's eight files."""
    exe = os.path.join(PKG, "mcgrid_shim")
    if not os.path.exists(exe):
        pytest.skip("driver shim not built")
    res = subprocess.run([exe, "--params", os.path.join(PKG, "driver", "input.params.example"), "--resident", "--calls", "4200",
                          "--out", str(tmp_path)], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stderr
    assert "-> 13390 loop iterations" in res.stdout                      # shipped parameters
    m = re.search(r"resident coupled loop: (\d+) iterations, (\d+) packets", res.stdout)
    assert m and int(m.group(1)) == 4200 and int(m.group(2)) == 4200 * 125000
    tail = "70w-80-500-400-0.030-0.030-0.060.dat"
    n = 80
    sizes = {"jmean-t": n ** 3, "rhokap-t": n ** 3, "temp-t": (n + 2) ** 3, "water-t": n ** 3, "tissue-t": n ** 3,
             "time-t-1-": n ** 3, "time-t-2-": n ** 3, "time-t-3-": n ** 3}
    for stem, cnt in sizes.items():
        f = tmp_path / (stem + tail)
        assert f.exists() and f.stat().st_size == 8 * cnt, stem
    temp = np.fromfile(tmp_path / ("temp-t" + tail), dtype="<f8").reshape((n + 2,) * 3, order="F")
    assert temp[0, 5, 5] == 5.0 and temp[5, 5, 0] == 25.0 and 100.0 <= temp.max() < 500.0   # boiling reached, no ablation yet
    water = np.fromfile(tmp_path / ("water-t" + tail), dtype="<f8")
    assert water.max() == 0.75 and water.min() < 0.75
    tissue = np.fromfile(tmp_path / ("tissue-t" + tail), dtype="<f8")
    t1 = np.fromfile(tmp_path / ("time-t-1-" + tail), dtype="<f8")
    assert (tissue >= 0.53).sum() == (t1 > 0).sum() > 0                  # every damaged voxel has its first threshold time
