"""CPU tier: the reference arm of bench.py prints ONE JSON line with the contract's keys (the GPU arm
needs a device; its line is checked for the same keys on the B200 by tests marked gpu)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e"}


def run_bench(*args):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True,
                         timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    return json.loads(lines[0])


def test_reference_arm_line():
    d = run_bench("--impl", "reference", "--steps", "2", "--warmup", "3")
    assert BASE_KEYS <= set(d) and d["impl"] == "reference"
    assert d["metric"] == "fakequant_fwd_bwd_GBps" and d["unit"] == "GB/s" and d["higher_is_better"] is True
    assert d["config"]["workload"] == "fakequant_fwd_bwd_channelwise_C512_inner512" and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    assert d["e2e"] == {"value": d["value"], "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["vs_baseline"] is None and d["dtype"] == "f32" and d["data"] == "synthetic"


@pytest.mark.gpu
def test_gpu_arm_line():
    d = run_bench("--steps", "5", "--warmup", "3", "--qat-steps", "3", "--qat-warmup", "3")
    assert BASE_KEYS | {"roofline", "cpu_baseline", "gpu_launches", "clocks"} <= set(d)
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert r["traffic"] is None or r["traffic"] > 0.5 * r["algorithmic_bytes_per_launch"]
    assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0 and 0 < d["e2e"]["value"] < d["value"]
    assert d["gpu_launches"] == 3 * d["steps"] and d["n_gpus"] == 1 and d["scaling"] == "weak"
    assert d["cpu_baseline"]["kind"] == "port" and 0 < d["cpu_baseline"]["value"] < d["e2e"]["value"]
    assert d["qat"]["unit"] == "images/s" and d["qat"]["value"] > 0 and d["qat"]["e2e"]["h2d_bytes_per_step"] > 0
    c = d["qat_cifar_paper"]                     # BASELINE configs[1]: per-channel LEARNED scales, batch 256
    assert "error" not in c and c["config"]["per_gpu_batch"] == 256 and c["value"] > 0
    assert c["config"]["scale_grad_mode"].startswith("threshold") and d["qat_cifar_paper_ste"]["config"]["scale_grad_mode"].startswith("none")
    assert d["qat_b128"]["config"]["per_gpu_batch"] == 128 and d["qat_b128"]["value"] > 0            # configs[4] per GPU
    assert d["qat_custom_loss"]["config"]["workload"].endswith("maxbin") and d["qat_custom_loss"]["value"] > 0   # configs[2]
    assert d["qat_mnist"]["value"] > 0 and d["mnist_cpu"]["nested"]["value"] > 0                      # configs[0], GPU and CPU
    assert d["roofline"]["traffic"] is None or "unchanged" in d["roofline"]["traffic_source"]
    assert 0 < d["e2e"]["frac_of_pcie_bound"] <= 1.2
    assert d["value"] > 3000, "fwd+bwd below half of the HBM roofline: something is wrong with the build"
    lay = d["layouts_64Mi"]                      # the other two traversal regimes at the headline size
    assert "error" not in lay and lay["interleaved_C512"]["frac_of_measured_peak"] > 0.8 and lay["per_tensor"]["frac_of_measured_peak"] > 0.8
    lt = d["loss_terms_64Mi"]                    # custom-loss reductions on the three layouts
    assert "error" not in lt and len(lt) == 6 and all(v["fwd_us"] > 0 and v["bwd_us"] > 0 for v in lt.values())
