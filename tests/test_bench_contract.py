"""CPU: the bench.py contract that does not need a GPU - the reference arm prints ONE JSON line with the agreed keys
(timed on the host cores through the oracle port), and the product arm refuses to run without a CUDA device."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(*args):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True,
                          cwd=ROOT, timeout=900, env={**os.environ, "CUDA_VISIBLE_DEVICES": ""})


def test_reference_arm_prints_one_json_line():
    r = run("--impl", "reference", "--steps", "1", "--warmup", "0")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "unet_denoising_steps_per_sec" and d["unit"] == "steps/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1 and d["value"] > 0
    assert d["config"]["workload"].startswith("FFHQ AF-LDM UNet2DModel") and d["config"]["global_batch"] == 16
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["vs_baseline"] is None                 # BASELINE.md publishes no number for this metric


def test_product_arm_has_no_cpu_fallback():
    r = run("--steps", "1")
    assert r.returncode != 0
    assert "no CPU fallback" in (r.stderr + r.stdout)
