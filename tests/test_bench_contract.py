"""bench.py contract pieces that run without a GPU: the reference arm (the oracle port of the
reference's CPU algorithm timed on the host cores) prints one JSON line with the agreed keys."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2",
                        "--warmup", "3", "--level", "5"], cwd=ROOT, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    assert line["impl"] == "reference" and line["metric"] == "MLUPS" and line["unit"] == "MLUPS"
    assert line["higher_is_better"] is True and line["value"] > 0 and line["n_gpus"] == 1
    assert line["steps"] == 2 and line["dtype"] == "f64" and line["data"] == "synthetic"
    assert "workload" in line["config"] and line["config"]["workload"].startswith("cfg2")
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == line["value"] and cb["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": "MLUPS", "h2d_bytes_per_step": 0,
                           "d2h_bytes_per_step": 0}
    assert line["gpu_launches"] == 0


def test_reference_arm_other_ranks_stay_silent():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                        "--steps", "1", "--level", "5"], cwd=ROOT, env=env, capture_output=True, text=True,
                       timeout=300)
    assert r.returncode == 0 and not [l for l in r.stdout.splitlines() if l.startswith("{")]


def test_traffic_capture_is_keyed_by_the_kernel_source(tmp_path):
    """roofline.traffic comes from a committed ncu capture; an entry only counts for the kernel
    source it was taken of (hash of sweep_kernel.cuh + collide.cuh + ...), so it cannot go stale
    silently when the kernel changes"""
    sys.path.insert(0, ROOT)
    import bench
    sha = bench.kernel_source_hash()
    p = tmp_path / "traffic.json"
    p.write_text(json.dumps([
        {"kernel": "sweepKernel<19,trt>", "cells": 256 ** 3, "traffic_gb": 6.36, "source": "x", "source_sha": sha},
        {"kernel": "sweepKernel<27,mrt>", "cells": 256 ** 3, "traffic_gb": 9.2, "source": "y", "source_sha": "0" * 16}]))
    t = bench.measured_traffic("sweepKernel<19,trt>", 256 ** 3, str(p))
    assert t and 6.0e9 < t["bytes"] < 6.6e9 and t["source_sha"] == sha
    assert bench.measured_traffic("sweepKernel<19,trt>", 64 ** 3, str(p)) is None
    assert bench.measured_traffic("sweepKernel<27,mrt>", 256 ** 3, str(p)) is None     # stale capture
    assert bench.BYTES_PER_LUP == {19: 380, 27: 540}
    committed = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(committed):
        for e in json.load(open(committed)):
            assert {"kernel", "cells", "traffic_gb", "source", "source_sha"} <= set(e)
