"""CPU: the bench contract that can be checked without a GPU — the reference arm (`bench.py --impl reference` times the
unmodified OptimisedUtilSSE from oracle/_ref on the host cores) prints one JSON line with the contract's keys, ranks other than
0 stay silent under a multi-rank launch, and our arm refuses to run without a CUDA device instead of falling back."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BENCH = os.path.join(ROOT, "bench.py")
need_ref = pytest.mark.skipif(not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "lib", "libref_driver.so")),
                              reason="oracle/_ref not built")


def run(args, env=None):
    e = dict(os.environ)
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        e.pop(k, None)
    e.update(env or {})
    return subprocess.run([sys.executable, BENCH] + args, capture_output=True, text=True, timeout=600, env=e, cwd=ROOT)


@need_ref
def test_reference_arm_prints_the_contract_line():
    r = run(["--impl", "reference", "--steps", "2", "--warmup", "1", "--ref-instances", "200"])
    assert r.returncode == 0, r.stderr
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "impl", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["metric"] == "skinned_Mvert_per_s" and d["unit"] == "Mvert/s"
    assert d["steps"] == 2 and d["warmup"] == 1 and d["n_gpus"] == 1 and d["higher_is_better"] is True
    assert d["vs_baseline"] is None and d["dtype"] == "f32" and d["scaling"] == "weak"
    assert "Sinbad" in d["config"]["workload"] and "model" not in d["config"]
    assert d["value"] > 0 and d["ms_per_step"] > 0
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


@need_ref
def test_reference_arm_is_silent_on_other_ranks():
    r = run(["--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1"], env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_our_arm_needs_a_gpu():
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    r = run(["--steps", "1", "--warmup", "1"])
    assert r.returncode != 0 and r.stdout.strip() == ""
