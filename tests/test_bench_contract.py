"""bench.py's output contract, as far as it can be checked without a GPU:
  * `--impl reference` falls back to the CPU oracle port when no GPU / reference build can run, and prints exactly
    ONE JSON line on stdout with the contract's keys;
  * the product arm does NOT fall back: without a CUDA device it must fail loudly."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SMALL = ["--steps", "1", "--warmup", "1", "--width", "96", "--height", "72", "--frames", "6"]


def _has_gpu():
    try:
        from rpg_open_remode_b200 import device_count
        return device_count() > 0
    except Exception:
        return False


def test_reference_arm_cpu_port_emits_one_json_line():
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--reference-cpu"] + SMALL,
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [l for l in res.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, res.stdout
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["unit"] == "frames/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and "workload" in d["config"]
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0 and d["vs_baseline"] is None


@pytest.mark.skipif(_has_gpu(), reason="a CUDA device is present: the product arm runs")
def test_product_arm_fails_loudly_without_a_gpu():
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + SMALL + ["--no-cpu-baseline"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert res.returncode != 0, "bench.py must not produce a number without the CUDA path"
    assert not any(l.strip().startswith("{") for l in res.stdout.splitlines()), res.stdout


def test_every_package_attribute_the_gpu_scripts_use_exists():
    """bench.py and the GPU-box tools only run where a B200 is: a misspelt or un-exported name must not cost a lease."""
    import re
    import rpg_open_remode_b200 as rmd
    files = [os.path.join(ROOT, "bench.py"), os.path.join(ROOT, "__graft_entry__.py")]
    for d in ("tools", os.path.join("tests", "perf")):
        files += [os.path.join(ROOT, d, f) for f in sorted(os.listdir(os.path.join(ROOT, d))) if f.endswith(".py")]
    missing = []
    for path in files:
        with open(path) as f:
            src = f.read()
        for name in set(re.findall(r"\brmd\.([A-Za-z_][A-Za-z_0-9]*)", src)):
            if not hasattr(rmd, name):
                missing.append((os.path.relpath(path, ROOT), name))
    assert not missing, missing
