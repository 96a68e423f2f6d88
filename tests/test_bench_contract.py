"""bench.py's reference arm on this box's host cores (no GPU involved): the JSON line the driver parses carries the
contract's keys, names BASELINE.json's metric and the same `config` the device arm prints."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _line(*args):
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    rows = [l for l in p.stdout.splitlines() if l.startswith("{")]
    assert len(rows) == 1, p.stdout[-2000:]                      # ONE json line
    return json.loads(rows[0])


def test_reference_arm_prints_the_contract_line():
    d = _line("--impl", "reference", "--steps", "1", "--warmup", "0")
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    assert "photon packets/s" in base["metric"]                  # BASELINE.json's metric, as one identifier + unit
    assert d["impl"] == "reference" and d["metric"] == "photon_packets_per_s" and d["unit"] == "packets/s"
    assert d["voxel_steps_per_s"] > 0                            # ... and its second quantity
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["dtype"] == "f64" and d["scaling"] == "strong"
    assert d["steps"] == 1 and d["warmup"] == 0 and d["n_gpus"] == 1 and d["value"] > 0
    assert d["config"]["workload"].startswith("skin200") and "model" not in d["config"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]


def test_both_arms_describe_the_same_workload():
    """config_block is the one place the workload is described; the device arm and the reference arm both call it."""
    sys.path.insert(0, ROOT)
    import bench

    src = open(os.path.join(ROOT, "bench.py")).read()
    assert src.count('"config": config_block(name, cfg, total)') == 2
    import tamc

    c = bench.config_block("skin200", tamc.configs.CONFIGS["skin200"], 10**9)
    assert c["packets_per_step"] == 10**9 and c["grid"] == "200^3" and "flushed" in c["l2"]
