"""csrc/tamc_math.cuh (the range-restricted fp64 log / sincospi / sqrt / reciprocal of the production kernels) compiled
for the host by plain g++ and measured against libm in long double: errors in ulps.  No GPU needed -- the device build
runs the same source with __fma_rn and the hardware reciprocal seeds."""
import json
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_elementary_functions_within_2_ulp(tmp_path):
    exe = str(tmp_path / "math_check")
    env = {k: v for k, v in os.environ.items() if k not in ("CC", "CXX")}
    subprocess.run(["g++", "-O2", "-std=c++17", "-x", "c++", "-o", exe, os.path.join(ROOT, "tools", "math_check.cpp")],
                   check=True, env=env)
    out = json.loads(subprocess.run([exe, "1500000"], check=True, capture_output=True, text=True).stdout)
    assert out["neglog_max_ulp"] < 1.0 and out["neglog_mean_ulp"] < 0.35
    assert out["sinpi_max_ulp"] < 2.0 and out["cospi_max_ulp"] < 2.0 and out["sinpi_mean_ulp"] < 0.4
    assert out["sqrt_max_ulp"] <= 0.5001 and out["rcp_max_ulp"] <= 1.0
