"""bench.py contract, as far as it can be checked without a GPU: the reference arm (`--impl reference`: the compiled
reference Simulator on the host cores) prints ONE JSON line with the keys the driver reads; the B200 arm refuses to run
without a CUDA device instead of falling back."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, env=None):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, timeout=600,
                          env=dict(os.environ, **(env or {})))


def test_reference_arm_prints_one_json_line():
    r = _run(["--impl", "reference", "--cpu-grid", "12", "--steps", "1", "--warmup", "1"])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "particle_updates_per_s" and d["unit"] == "particle-updates/s"
    assert d["value"] > 0 and d["higher_is_better"] is True and d["vs_baseline"] is None and d["data"] == "synthetic"
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"]


def test_reference_arm_other_ranks_exit_quietly():
    r = _run(["--impl", "reference", "--gpus", "2", "--cpu-grid", "12", "--steps", "1", "--warmup", "1"], env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_b200_arm_has_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    r = _run(["--steps", "1", "--warmup", "1", "--grid", "16"])
    assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout)


@pytest.mark.parametrize("name", ["r2_bench_n1_final.json", "r2_bench_slab_n2.json", "r2_bench_slab_n8.json"])
def test_committed_b200_lines_carry_the_contract(name):
    """The bench lines committed under profiles/ (written by bench.py on B200 boxes) carry every key the driver and the judge
    read: the base contract, `roofline`, `e2e`, `clocks`, `gpu_launches`, the correctness gate, and -- at N = 1 -- `cpu_baseline`."""
    d = json.load(open(os.path.join(ROOT, "profiles", name)))
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
              "config", "roofline", "e2e", "clocks", "gpu_launches", "checks"):
        assert k in d, k
    assert d["metric"] == "particle_updates_per_s" and d["unit"] == "particle-updates/s" and d["higher_is_better"] is True
    assert d["scaling"] == "strong" and d["vs_baseline"] is None and d["data"] == "synthetic" and "workload" in d["config"]
    assert abs(d["value"] - d["config"]["particles_total"] / (d["ms_per_step"] * 1e-3)) <= 1e-6 * d["value"]
    r = d["roofline"]
    assert r["kernel"] == "p2g_kernel" and r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    e = d["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and e["value"] < d["value"]
    assert d["clocks"]["sm_mhz"] > 0 and not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    assert d["gpu_launches"] > 0 and d["checks"]["ok"] is True
    if d["n_gpus"] == 1:
        c = d["cpu_baseline"]
        assert c["kind"] == "reference" and c["cores"] >= 1 and c["value"] > 0 and c["sample"]
    else:
        assert d["checks"]["slab_verify"]["ok"] is True and d["exchange_wait"] is not None
