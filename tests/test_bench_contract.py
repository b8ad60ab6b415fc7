"""CPU: the parts of bench.py that run without a GPU keep the contract of the task statement -- the reference arm prints
one JSON line with the agreed keys, and the product arm fails loudly (no CPU fallback) when there is no CUDA device."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_bench(*args, timeout=600):
    env = dict(os.environ, OMP_NUM_THREADS="1")
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + list(args), cwd=ROOT, env=env, timeout=timeout,
                          capture_output=True, text=True)


def test_reference_arm_prints_the_contract_line():
    out = run_bench("--impl", "reference", "--steps", "1", "--warmup", "0", "--workload", "cfg2")
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().split("\n")[-1])
    assert line["impl"] == "reference" and line["metric"] == "bootstrap ML fits/sec" and line["unit"] == "replicate-fits/s"
    assert line["higher_is_better"] is True and line["value"] > 0 and line["n_gpus"] == 1 and line["steps"] == 1
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1 and line["cpu_baseline"]["value"] == line["value"]
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["config"]["workload"].startswith("cfg2:") and "model" not in line["config"]
    assert line["gpu_launches"] == 0 and line["vs_baseline"] is None


def test_product_arm_has_no_cpu_fallback(lib_built):
    from traversome_b200 import _cabi
    if _cabi.device_count() > 0:
        pytest.skip("a GPU is present")
    out = run_bench("--steps", "1", "--warmup", "0", "--no-cpu-baseline", "--no-selection")
    assert out.returncode != 0
    assert "value" not in out.stdout                      # no bench line without a device
