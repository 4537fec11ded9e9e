"""bench.py contract checks that need no GPU: the CPU reference arm prints one well-formed JSON line (rank 0 only under a
launcher), and the engine arm refuses to run without a CUDA device instead of falling back to anything."""
import json
import os
import subprocess
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BENCH = os.path.join(ROOT, "bench.py")


def _run(args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, BENCH] + args, capture_output=True, text=True, timeout=600, env=e)


def test_reference_arm_line():
    r = _run(["--impl", "reference", "--steps", "1", "--warmup", "1", "--cpu-sample", "64"])
    assert r.returncode == 0, r.stderr
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "bn254_verifies_per_sec" and d["unit"] == "verifies/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["steps"] == 1 and d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"] == d["e2e"]["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["vs_baseline"] is None


def test_reference_arm_other_ranks_stay_silent():
    r = _run(["--impl", "reference", "--steps", "1", "--warmup", "0", "--cpu-sample", "64", "--gpus", "2"], env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_engine_arm_needs_a_gpu():
    if torch.cuda.is_available():
        return  # on a GPU box the engine arm is exercised by the driver itself
    r = _run(["--steps", "1", "--warmup", "0"])
    assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout)
