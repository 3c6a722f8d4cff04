"""bench.py contract checks that need no GPU: the reference arm runs on host cores and prints
one JSON line with the agreed keys; under a multi-rank launch only rank 0 prints."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(env_extra=None, args=()):
    env = dict(os.environ)
    env.update(env_extra or {})
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1",
                          "--ref-particles", "2000", *args], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True,
                         env=env, timeout=300)
    assert res.returncode == 0, res.stderr[-2000:]
    return res.stdout.strip()


def test_reference_arm_prints_contract_line():
    out = _run()
    lines = [ln for ln in out.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"):
        assert key in d, key
    assert d["impl"] == "reference" and d["unit"] == "rays/s" and d["higher_is_better"] is True
    assert d["vs_baseline"] is None and d["data"] == "synthetic" and d["dtype"] == "f64"
    assert d["steps"] == 2 and d["value"] > 0
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "Spielberg_map" in d["config"]["workload"]
    assert "model" not in d["config"]


def test_reference_arm_other_ranks_stay_silent():
    out = _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}, ("--gpus", "2"))
    assert out == ""
