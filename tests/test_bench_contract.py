"""bench.py contract checks that need no GPU: the reference arm (CPU oracle) prints exactly one JSON line with the
keys the driver reads, and the B200 arm refuses to run without a CUDA device (no CPU fallback)."""
import json
import os
import sys as _sys
_sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, cwd=ROOT, env=e,
                          timeout=600)


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    r = _run(["--impl", "reference", "--config", "forest", "--steps", "3", "--warmup", "1"])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["n_gpus"] == 1
    for k in ("metric", "value", "unit", "steps", "warmup", "ms_per_step", "scaling", "vs_baseline", "dtype", "data", "config",
              "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["value"] > 0 and d["config"]["workload"].startswith("params/forest_best.py")
    # the unmodified reference when a checkout is present (build container: /root/reference; GPU box: baseline/_ref),
    # else the oracle port
    from oracle.reference_access import find_reference
    assert d["cpu_baseline"]["kind"] == ("reference" if find_reference() else "port")
    assert d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    # the driver divides the two arms' values only when unit, steps and warm-up agree: the arm honours its arguments
    assert d["unit"] == "HVP/s" and d["steps"] == 3 and d["warmup"] == 1
    import re
    src = open(os.path.join(ROOT, "bench.py")).read()
    assert re.search(r'^UNIT = "HVP/s"$', src, flags=re.M) and src.count('"unit": UNIT') >= 2
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_exit_without_work():
    r = _run(["--impl", "reference", "--config", "forest", "--steps", "1", "--warmup", "0", "--gpus", "2"], env={"RANK": "1", "WORLD_SIZE": "2"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_b200_arm_fails_loudly_without_a_cuda_device():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a CUDA device is present")
    r = _run(["--steps", "1", "--warmup", "0", "--no-cpu", "--no-vec", "--no-reg"])
    assert r.returncode != 0
    assert "no CPU fallback" in r.stderr and r.stdout.strip() == ""
