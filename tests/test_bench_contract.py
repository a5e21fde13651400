"""bench.py's CPU-runnable leg: `--impl reference` must run without a GPU (it times the reference's
CPU implementation of the path on the host cores) and print ONE JSON line with the contract's keys.
The GPU arm's line is checked on the GPU tier."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
             "dtype", "data", "config", "e2e", "cpu_baseline"}


def run(*args):
    out = subprocess.check_output([sys.executable, os.path.join(ROOT, "bench.py"), *args], text=True, cwd=ROOT, stderr=subprocess.DEVNULL)
    lines = [l for l in out.splitlines() if l.strip()]
    assert len(lines) == 1, "exactly one JSON line"
    return json.loads(lines[0])


def test_reference_arm_runs_on_the_host_cores():
    d = run("--impl", "reference", "--workload", "cornell_2x2x2", "--steps", "2", "--warmup", "1")
    assert BASE_KEYS <= set(d) and d["impl"] == "reference"
    assert d["metric"] == "probe_rays_per_s" and d["unit"] == "probe-rays/s" and d["higher_is_better"] is True
    assert d["config"]["workload"] == "cornell_2x2x2" and d["config"]["probe_rays"] == 512
    assert d["value"] > 0 and d["gpu_launches"] == 0
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["value"] == d["value"] and cb["cores"] >= 1 and cb["kind"] in ("reference", "port") and cb["sample"]
    from oracle import ref

    # the reference's own shader text is the CPU arm wherever it can run the workload
    assert cb["kind"] == ("reference" if ref.available() else "port")


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "cornell_2x2x2", "--gpus", "2"],
                       capture_output=True, text=True, cwd=ROOT, env=env)
    assert p.returncode == 0 and p.stdout.strip() == ""


@pytest.mark.gpu
def test_gpu_arm_line_has_the_contract_keys():
    d = run("--workload", "cave_64", "--steps", "3", "--warmup", "3", "--no-cpu-baseline")
    assert BASE_KEYS <= set(d)
    assert {"roofline", "clocks", "gpu_launches", "fps"} <= set(d)
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12
    assert d["gpu_launches"] == d["steps"] and d["n_gpus"] == 1
    e = d["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] == 1024 * 64 * 4
    assert e["value"] != d["value"]
    # one GPU: one frame at a time unless asked for; the pipeline behind `value` is named
    assert d["config"]["frames_in_flight"] == 1 and d["pipelines"]["value_is"] == "one_frame_at_a_time"
    assert d["pipelines"]["two_frames_in_flight"] is None
    d2 = run("--workload", "cave_64", "--steps", "3", "--warmup", "3", "--no-cpu-baseline", "--no-ncu", "--frames-in-flight", "2")
    assert d2["config"]["frames_in_flight"] == 2 and d2["pipelines"]["value_is"] == "two_frames_in_flight"
    assert d2["gpu_launches"] == d2["steps"] and d2["ms_per_step"] == d2["pipelines"]["two_frames_in_flight"]["ms_per_step"]
