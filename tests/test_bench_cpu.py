"""bench.py contract checks that need no GPU: the reference arm prints one JSON
line with the agreed keys; the B200 arm refuses to run without CUDA."""
import json
import os
import subprocess
import sys

import pytest
import torch

import conftest

BENCH = os.path.join(conftest.ROOT, "bench.py")


def test_reference_arm_prints_one_json_line():
    out = subprocess.run([sys.executable, BENCH, "--impl", "reference", "--config", "cfg1", "--steps", "2",
                          "--warmup", "1", "--cpu-sample", "8"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "patterns/s" and d["value"] > 0
    assert d["metric"] == "spai_patterns_scored_per_s" and d["higher_is_better"] is True
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    assert d["steps"] == 2 and d["warmup"] == 1 and "workload" in d["config"]
    # both arms print the same `config` dict (bench.config_dict): static facts of the workload only
    assert set(d["config"]) == {"workload", "mode", "n", "num_edges", "batch_per_gpu", "alpha", "max_deleted_fraction",
                                "trajectories", "parallelism", "l2", "timing"}
    it = d["reference_itself"]
    assert it is None or "error" in it or (it["as_is_patterns_per_s"] > 0 and it["minus_gc_patterns_per_s"] > 0)


def test_reference_arm_non_zero_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, BENCH, "--impl", "reference", "--gpus", "2", "--config", "cfg1"],
                         capture_output=True, text=True, timeout=120, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


@pytest.mark.skipif(torch.cuda.is_available(), reason="CUDA present")
def test_b200_arm_fails_loudly_without_cuda():
    out = subprocess.run([sys.executable, BENCH, "--config", "cfg1", "--steps", "1"], capture_output=True,
                         text=True, timeout=300)
    assert out.returncode != 0
    assert "no CPU fallback" in out.stderr or "CUDA" in out.stderr
