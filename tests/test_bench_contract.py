"""bench.py's reference arm (no GPU): the JSON line the driver parses, and the rank rule.

`bench.py --impl reference` times the UNMODIFIED reference (oracle/_ref) through its own
`tesseroid.gz(..., njobs=cores)` on a bounded sample of the C4 workload; under torchrun only rank 0
runs and prints, the other ranks exit 0 without work."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_bench(extra_env, *args):
    env = dict(os.environ)
    env.update(extra_env)
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + list(args), env=env,
                          capture_output=True, text=True, timeout=600)


def test_reference_arm_prints_the_contract_line():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import build_ref
    if not build_ref.available():
        pytest.skip("oracle/_ref not built")
    out = run_bench({"RANK": "0"}, "--impl", "reference", "--gpus", "1", "--steps", "1",
                    "--warmup", "0")
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    line = json.loads(lines[0])
    assert line["impl"] == "reference"
    assert line["metric"] == "tesseroid-point evals/sec (gz, exp density)" and line["unit"] == "evals/s"
    assert line["higher_is_better"] is True and line["n_gpus"] == 1 and line["steps"] == 1
    assert line["value"] > 0 and line["ms_per_step"] > 0
    assert line["config"]["workload"].startswith("C4 synthetic regional model")
    base = line["cpu_baseline"]
    assert base["kind"] == "reference" and base["cores"] == (os.cpu_count() or 1)
    assert base["value"] == line["value"] and "random tesseroids" in base["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": "evals/s", "h2d_bytes_per_step": 0,
                           "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_do_nothing():
    out = run_bench({"RANK": "1", "WORLD_SIZE": "2"}, "--impl", "reference", "--gpus", "2",
                    "--steps", "1", "--warmup", "0")
    assert out.returncode == 0, out.stderr[-2000:]
    assert out.stdout.strip() == ""


def test_gpu_arm_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    out = run_bench({}, "--steps", "1", "--warmup", "3", "--no-cpu", "--legs", "none")
    assert out.returncode != 0
    assert "no CPU fallback" in out.stderr
    assert not [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
