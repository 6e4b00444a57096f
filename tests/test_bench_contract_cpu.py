"""CPU: the parts of bench.py that do not need a GPU -- the work model of SURVEY §8d, the CPU-baseline extrapolation
and the committed result lines under profiles/ carrying every key of the bench contract."""
import json
import os

import pytest

import bench

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_work_model_matches_survey_numbers():
    # C2 (B=2, k=25) W_ref ~ 97 TFLOP, W_min ~ 89 TFLOP (3B LoRA-off samples; 2 distinct prompts -> 87.2)
    assert abs(bench.w_ref_tflop(2, 25) - 97.22) < 0.05
    assert abs(bench.w_min_tflop(2, 25, distinct_nograd=3) - 88.81) < 0.05
    assert abs(bench.w_min_tflop(2, 25) - 87.20) < 0.05
    assert bench.w_min_tflop(2, 25) < bench.w_ref_tflop(2, 25)
    # one iteration = (k+4) CFG forwards + backward of 2B samples at 1.23 F
    sec_cfg_fwd = 4.0                                   # 2 samples
    per_sample = sec_cfg_fwd / 2
    t_iter = 2 * 2 * 29 * per_sample + 2 * 2 * 1.23 * per_sample
    assert abs(bench.cpu_latents_per_s(sec_cfg_fwd, 2, 25) - 2 / t_iter) < 1e-12


REQUIRED = ["metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
            "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "roofline", "clocks"]


@pytest.mark.parametrize("name", ["r1_final_bench.json", "r1_bench_2gpu.json"])
def test_committed_bench_lines_follow_the_contract(name):
    line = json.load(open(os.path.join(ROOT, "profiles", name)))
    for k in REQUIRED:
        assert k in line, k
    assert line["metric"] == "leco_train_latents_per_sec" and line["unit"] == "latents/s" and line["higher_is_better"]
    assert "workload" in line["config"] and "model" not in line["config"]
    assert set(line["e2e"]) >= {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"}
    assert line["e2e"]["h2d_bytes_per_step"] > 0 and line["e2e"]["d2h_bytes_per_step"] > 0
    r = line["roofline"]
    assert set(r) >= {"bound", "achieved", "peak", "unit", "frac", "traffic"} and r["bound"] in ("hbm", "tensor")
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and r["traffic"] is not None
    assert line["gpu_launches"] > 1000 and line["warmup"] >= 3
    assert not set(line["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    if line["n_gpus"] == 1:
        assert set(line["cpu_baseline"]) >= {"value", "unit", "cores", "kind", "sample"}
        assert abs(line["value"] - line["config"]["global_batch"] * 1000.0 / line["ms_per_step"]) < 1e-6


def test_committed_reference_arm_line():
    line = json.load(open(os.path.join(ROOT, "profiles", "r1_final_bench_reference_arm.json")))
    assert line["impl"] == "reference" and line["metric"] == "leco_train_latents_per_sec"
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["value"] == line["value"]
