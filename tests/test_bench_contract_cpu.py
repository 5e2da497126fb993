"""CPU suite: bench.py's reference arm (the reference's CPU algorithm on the host cores) runs without a GPU and prints
exactly one JSON line with the contract's keys; the GPU arm refuses to run without a device (no CPU fallback)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1",
                        "--width", "64", "--skip-ref-binary"], capture_output=True, text=True, timeout=300,
                       env=dict(os.environ, OMP_NUM_THREADS="1"))       # what torchrun exports to every rank
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "nn_pairs_per_sec" and d["unit"] == "pairs/s" and d["higher_is_better"] is True
    assert d["steps"] == 2 and d["warmup"] == 1 and d["value"] > 0 and d["ms_per_step"] > 0
    assert d["cpu_baseline"]["kind"] in ("port", "reference") and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"]
    assert d["cpu_baseline"]["cores"] == (os.cpu_count() or 1), "the CPU arm must use every host core whatever OMP_NUM_THREADS says"


def test_reference_arm_every_config():
    """--config 2, 3, 5 of the CPU arm: one contract line each, naming the configuration it measured."""
    for cfg, extra in ((2, ["--width", "48"]), (3, ["--width", "48"]), (5, ["--batch", "16"])):
        r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", str(cfg), "--steps", "1", "--warmup", "0",
                            "--skip-ref-binary"] + extra, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        d = json.loads(r.stdout.strip().splitlines()[-1])
        assert d["impl"] == "reference" and d["config"]["config_id"] == cfg and d["value"] > 0 and d["unit"] == "pairs/s"
        assert ("batched" in d["config"]["workload"]) == (cfg == 5)
        assert ("point-to-plane" in d["config"]["workload"]) == (cfg == 3)


def test_gpu_arm_fails_loudly_without_a_device():
    import ctypes
    n = ctypes.c_int(0)
    sys.path.insert(0, os.path.join(ROOT, "fast-point-cloud-registration-with-gpus_b200", "python"))
    import icp_b200 as ib
    if ib.lib.icpb_device_count(ctypes.byref(n)) == 0 and n.value > 0:
        import pytest
        pytest.skip("a GPU is present")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "0", "--width", "32"],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode != 0 and "no CPU path" in (r.stderr + r.stdout)
