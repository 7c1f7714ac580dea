"""The committed bench lines (profiles/r1_bench*_line.json, written by bench.py on a B200) carry every key of the
bench contract; bench.py's command line parses without a GPU.  Guards the JSON shape the driver reads."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _line(name):
    with open(os.path.join(ROOT, "profiles", name)) as f:
        return json.loads(f.read().strip().splitlines()[-1])


def _base_keys(d):
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "e2e"):
        assert k in d, k
    # the metric key is a slug of BASELINE.json's wording ("ρ·RHS evals/sec at 10 qubits (batch) …"); newer lines carry the
    # wording itself in config.baseline_metric
    assert d["metric"] == "rho_rhs_evals_per_sec_10q_ame_batch64" and d["unit"] == "evals/s"
    if "baseline_metric" in d["config"]:
        assert d["config"]["baseline_metric"] == json.load(open(os.path.join(ROOT, "BASELINE.json")))["metric"]
    assert d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None
    assert d["data"] == "synthetic" and "workload" in d["config"] and "model" not in d["config"]
    for k in ("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"):
        assert k in d["e2e"], k


@pytest.mark.parametrize("name,n", [("r1_bench_line.json", 1), ("r1_bench_2gpu_line.json", 2), ("r1_bench_4gpu_line.json", 4),
                                    ("r1_bench_8gpu_line.json", 8)])
def test_own_arm_line(name, n):
    d = _line(name)
    _base_keys(d)
    assert d["n_gpus"] == n and d["warmup"] >= 3 and d["value"] > 0
    assert d["gpu_launches"] > 0
    assert set(("sm_mhz", "sm_max_mhz", "reasons")) <= set(d["clocks"])
    assert not any(r in ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown") for r in d["clocks"]["reasons"])
    # inputs really cross PCIe in the end-to-end number, and it is not a copy of the device-timed value
    per_gpu = 64 * 1024 * 1024 * 16
    assert d["e2e"]["h2d_bytes_per_step"] >= per_gpu and d["e2e"]["d2h_bytes_per_step"] >= per_gpu
    assert d["e2e"]["value"] != d["value"]
    if n == 1:
        r = d["roofline"]
        for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
            assert k in r, k
        assert r["bound"] == "tensor" and r["unit"] == "TFLOP/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
        assert 0.5 < r["frac_executed"] <= 1.0  # pipe utilisation is a real fraction even when frac (algorithmic) exceeds 1
        assert r["general_complex_eigenvectors"]["rel_diff_real_vs_general_product"] < 1e-12
        c = d["cpu_baseline"]
        for k in ("value", "unit", "cores", "kind", "sample"):
            assert k in c, k
        assert c["kind"] in ("port", "reference")
        # the step is weakly scaled: per-GPU work is the same at every N
        assert d["config"]["batch_per_gpu"] == 64


def test_reference_arm_line():
    d = _line("r1_bench_reference_line.json")
    _base_keys(d)
    assert d["impl"] == "reference" and d["n_gpus"] == 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    assert d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["cores"] >= 1
    own = _line("r1_bench_line.json")
    assert d["metric"] == own["metric"] and d["unit"] == own["unit"] and d["config"]["workload"] == own["config"]["workload"]


def test_bench_cli_parses_without_gpu():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--help"], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and "--gpus" in out.stdout and "--impl" in out.stdout


@pytest.mark.parametrize("name,n", [("r2_bench_line.json", 1), ("r2_bench_2gpu_line.json", 2)])
def test_round2_line(name, n):
    """Round-2 shape of the line: the roofline is split into the structure row (what the timed step runs; frac is a real
    fraction of the measured peak) and the contract row (general complex product); `traffic` is null with the profile it
    comes from named; like-for-like e2e, host link probe and (N > 1) the library's NCCL reduce are present."""
    d = _line(name)
    _base_keys(d)
    assert d["n_gpus"] == n and d["warmup"] >= 3 and d["value"] > 0 and d["gpu_launches"] > 0
    assert not any(r in ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown") for r in d["clocks"]["reasons"])
    per_gpu = 64 * 1024 * 1024 * 16
    assert d["e2e"]["h2d_bytes_per_step"] >= per_gpu and d["e2e"]["d2h_bytes_per_step"] >= per_gpu and d["e2e"]["value"] != d["value"]
    lfl = d["like_for_like_e2e"]
    assert lfl["value"] > 0 and lfl["plans_built"] >= lfl["steps"] and lfl["rel_diff_vs_host_lapack_plan_at_s0"] < 1e-10
    link = d["host_link"]
    assert 0 < link["e2e_fraction_of_link"] <= 1.0
    if n == 1:
        r = d["roofline"]
        for k in ("bound", "achieved", "peak", "unit", "frac", "traffic", "row", "contract_row"):
            assert k in r, k
        assert r["bound"] == "tensor" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
        assert 0.5 < r["frac"] <= 1.0                      # the structure row is a fraction of the pipe, never above 1
        assert r["traffic"] is None and "from_profile_not_this_run" in r["traffic_profile"]
        cr = r["contract_row"]
        assert abs(cr["frac"] - cr["achieved"] / cr["peak"]) < 1e-9 and 0.5 < cr["frac_executed"] <= 1.0
        assert cr["rel_diff_vs_structure_row_result"] < 1e-12
        c = d["cpu_baseline"]
        assert c["kind"] == "port" and c["cores"] >= 1 and set(c["literal_davies_loop"]["per_level"]) == {"16", "32", "64"}
        sec = d["secondary"]
        for k in ("C2_dephasing_diagL", "C4_traj_xor", "C4_generic_kernel_same_operators", "C4_random_pattern", "C5_per_gpu"):
            assert k in sec and sec[k]["roofline"]["frac"] > 0, k
        assert sec["C4_traj_xor"]["roofline"]["bound"] == "hbm" and sec["C4_traj_xor"]["roofline"]["frac"] > 0.7
    else:
        o = d["obs_allreduce"]
        assert o["calls"] >= 20 and o["warmup_calls"] >= 5 and 0 < o["allreduce_ms_median"] < 1.0 and o["rel_diff_vs_torch_path"] < 1e-12
