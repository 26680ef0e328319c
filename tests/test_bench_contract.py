"""bench.py contract checks that need no GPU: the reference arm prints ONE JSON line with the agreed keys, and the
product arm refuses to run without a CUDA device (there is no CPU fallback to measure)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1", "--cpu-n", "12"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "elements/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["n_gpus"] == 1 and d["steps"] == 1 and d["gpu_launches"] == 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "sample" in d["config"]


def test_product_arm_needs_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=600)
    assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout)
    assert not [l for l in r.stdout.splitlines() if l.startswith("{")]


def test_time_stats_json_is_readable_by_the_reference_scripts(tmp_path):
    """arcanefem_b200/timestats.py writes what modules/testlab/benchmarking/get_stats_from_json.py reads: the lookup below
    restates that script's find_key / Cumulative / cacheWarming arithmetic (:5-21, :78-100)."""
    import json
    from arcanefem_b200.timestats import write_time_stats

    def find_key(data, target_key):
        if isinstance(data, dict):
            for key, value in data.items():
                if key == target_key:
                    return value
                result = find_key(value, target_key)
                if result is not None:
                    return result
        return None

    p = tmp_path / "time_stats.json"
    write_time_stats(str(p), 10, 1, 3, 17 ** 3, 12 * 16 * 16, 6 * 16 ** 3, {"csr-gpu": (1e-3, 2e-3), "nwcsr": (1e-4, 4e-4)})
    obj = json.load(open(p))
    for k in ("cacheWarming", "nbParallelInstance", "meshDim", "nbNode", "nbBoundaryElement", "nbElement", "acceleratorRuntime"):
        assert k in obj
    cw = obj["cacheWarming"]
    for name, (bm, ac) in (("Csr_Gpu", (1e-3, 2e-3)), ("CsrNodeWise", (1e-4, 4e-4))):
        node = find_key(obj, "AssembleBilinearOperator_" + name)
        assert abs(float(node["Cumulative"].split(" ")[0]) / (cw - 1) - (bm + ac)) < 1e-12
        assert abs(float(find_key(node, "BuildMatrix")["Cumulative"].split(" ")[0]) / (cw - 1) - bm) < 1e-12
        assert abs(float(find_key(node, "AddAndCompute")["Cumulative"].split(" ")[0]) / (cw - 1) - ac) < 1e-12
