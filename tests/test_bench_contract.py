"""bench.py's driver contract, CPU tier: the reference arm (`--impl reference`) runs here without a GPU and prints ONE JSON
line with the keys the driver reads; the source of our arm names every contract key."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.strip().splitlines() if ln.startswith("{")]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "train_rays_per_s" and d["unit"] == "rays/s"
    assert d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None
    assert d["value"] > 0 and d["steps"] == 1 and d["n_gpus"] == 1
    assert d["config"]["grid"] == 160 and d["config"]["rays_per_gpu"] == 8192 and "workload" in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["sample"] and cb["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # the arm must not map the product library into its process (it times the reference formulation only)
    assert "libposeprobe_b200" not in r.stderr


def test_our_arm_names_every_contract_key():
    src = open(os.path.join(ROOT, "bench.py")).read()
    for key in ('"metric"', '"value"', '"unit"', '"n_gpus"', '"steps"', '"warmup"', '"ms_per_step"', '"higher_is_better"',
                '"scaling"', '"vs_baseline"', '"dtype"', '"data"', '"config"', '"roofline"', '"bound"', '"achieved"',
                '"peak"', '"frac"', '"traffic"', '"e2e"', '"h2d_bytes_per_step"', '"d2h_bytes_per_step"',
                '"gpu_launches"', '"clocks"', '"cpu_baseline"', '"kind"', '"cores"', '"sample"'):
        assert key in src, key
