"""bench.py's contract where it can run without a GPU: the reference arm (oracle/_ref or the port) prints ONE JSON line
with the agreed keys, and the product arm refuses to run without CUDA instead of falling back to the CPU."""
import json
import os
import subprocess
import sys

import pytest

from xm_helpers import ROOT

torch = pytest.importorskip("torch")


def run_bench(*args, timeout=300):
    env = dict(os.environ)
    env.pop("RANK", None)
    env.pop("WORLD_SIZE", None)
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=timeout, env=env)


def test_reference_arm_json_line():
    out = run_bench("--impl", "reference", "--steps", "1", "--warmup", "1", "--events", "20000", "--ref-workers", "2")
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "events/sec" and d["unit"] == "events/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["steps"] == 1 and d["warmup"] == 1
    assert d["config"]["workload"].startswith("synthetic Poisson stream")
    from oracle import ref_chain

    # the reference's own modules (oracle/_ref) when they were built, the NumPy restatement otherwise
    assert d["cpu_baseline"]["kind"] == ("reference" if ref_chain.available() else "port")
    assert d["cpu_baseline"]["cores"] == 2 and d["cpu_baseline"]["value"] == d["value"]
    # both arms print the same `config` (the driver compares them)
    sys.path.insert(0, ROOT)
    import argparse

    import bench

    assert d["config"] == bench.make_config("5m", argparse.Namespace(events=20000))
    assert d["e2e"] == {"value": d["value"], "unit": "events/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=120, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


@pytest.mark.skipif(torch.cuda.is_available(), reason="only meaningful on a machine without CUDA")
def test_product_arm_has_no_cpu_fallback():
    out = run_bench("--steps", "1", timeout=120)
    assert out.returncode != 0
    assert "no CPU path" in (out.stderr + out.stdout)


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under x-maps_b200/ may import or execute it."""
    import re

    pkg = os.path.join(ROOT, "x-maps_b200")
    offenders = []
    for base, _, files in os.walk(pkg):
        for name in files:
            if not name.endswith((".py", ".cu", ".cuh", ".h")):
                continue
            text = open(os.path.join(base, name), encoding="utf-8", errors="replace").read()
            if re.search(r"^\s*(from|import)\s+oracle\b", text, re.M) or "xmaps_oracle" in text:
                offenders.append(os.path.join(base, name))
    assert not offenders, offenders
