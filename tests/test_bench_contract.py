"""bench.py's one-line JSON contract (the driver parses it): keys, units and the arithmetic that ties
them together.  The reference arm runs on the CPU (non-gpu test, reduced sample); the GPU arm is
checked on a small mesh."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, env=e,
                       timeout=900, cwd=ROOT)
    assert p.returncode == 0, p.stdout + p.stderr
    return json.loads(p.stdout.strip().splitlines()[-1])


def test_reference_arm_line():
    j = _run(["--impl", "reference", "--steps", "1", "--warmup", "0"], {"BENCH_CPU_SAMPLE_N": "96"})
    assert j["impl"] == "reference" and j["metric"] == "GLL DOF-updates/sec" and j["unit"] == "DOF-updates/s"
    assert j["higher_is_better"] is True and j["value"] > 1e6
    cb = j["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] == 1 and cb["value"] == j["value"] and "96x96" in cb["sample"]
    assert j["e2e"] == {"value": j["value"], "unit": j["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in j["config"] and "model" not in j["config"]


@pytest.mark.gpu
def test_gpu_arm_line():
    j = _run(["--nx", "512", "--nz", "512", "--steps", "5", "--warmup", "3", "--generic-n", "256", "--no-configs"],
             {"BENCH_CPU_SAMPLE_N": "96"})
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks"):
        assert k in j, k
    assert j["n_gpus"] == 1 and j["steps"] == 5 and j["dtype"] == "f64" and j["data"] == "synthetic" and j["vs_baseline"] is None
    ndof = j["config"]["npoin_per_gpu"] * 2
    assert abs(j["value"] - ndof * 5 / (j["ms_per_step"] * 5e-3)) <= 1e-6 * j["value"]
    r = j["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12
    assert abs(r["achieved"] - r["algorithmic_bytes_per_dof"] * r["dofs_per_launch"] / (r["ms_per_launch"] * 1e-3) / 1e9) <= 1e-6 * r["achieved"]
    assert r["algorithmic_bytes_per_dof"] == 48.5      # (lambda, mu) + d, v, rmass/2 in; v, d_next out
    assert j["e2e"]["h2d_bytes_per_step"] > 0 and j["e2e"]["d2h_bytes_per_step"] > 0 and j["e2e"]["value"] > 0
    assert j["gpu_launches"] >= 5 and j["cpu_baseline"]["cores"] == 1
    assert set(j["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
    assert j["gpu_launches"] == 6 * 5          # six launches per fused step
    g = j["generic_route"]
    assert g["kernel_route"] == "strip kernel" and 0.8 < g["generic_over_builder"] < 1.25
    assert set(j["ms_per_step_by_phase"]) >= {"element_force", "boundary_conditions", "outputs"}
    assert j["cpu_baseline"]["parity_build_value"] > 0


@pytest.mark.gpu
@pytest.mark.parametrize("name,scale", [("testsh", 2), ("lamb", 2), ("tpv3", 2), ("ratestate", 1)])
def test_reference_config_records(name, scale):
    """bench.py --config: BASELINE.json configs[0..3] through the host program (small scale here)"""
    j = _run(["--config", name, "--config-scale", str(scale), "--steps", "20"])
    assert "error" not in j, j
    assert j["config"] == name and j["value"] > 0 and j["roofline"]["frac"] > 0
    assert j["kelvin_voigt"] == (name == "tpv3")
    assert abs(j["value"] - j["npoin"] * j["ndof"] / (j["ms_per_step"] * 1e-3)) <= 1e-6 * j["value"]
