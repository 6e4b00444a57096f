"""CPU: the parts of bench.py that do not need a GPU -- the work model of SURVEY §8d, the CPU-baseline extrapolation
and the committed result lines under profiles/ carrying every key of the bench contract."""
import json
import os

import pytest

import bench

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_work_model_matches_survey_numbers():
    # C2 (B=2, k=25) W_ref ~ 97 TFLOP, W_min ~ 89 TFLOP (3B LoRA-off samples; 2 distinct prompts -> 87.2)
    c2 = bench.CONFIGS["sd21"]
    assert abs(bench.w_ref_tflop(c2, 25) - 97.22) < 0.05
    assert abs(bench.w_min_tflop(c2, 25, distinct_nograd=3) - 88.81) < 0.05
    assert abs(bench.w_min_tflop(c2, 25) - 87.20) < 0.05
    assert bench.w_min_tflop(c2, 25) < bench.w_ref_tflop(c2, 25)
    # C3 / C5-per-rank (B=4) W_ref ~ 194 TFLOP, C4 (SDXL) ~ 816 TFLOP
    assert abs(bench.w_ref_tflop(bench.CONFIGS["sd21_b4"], 25) - 194.4) < 0.5
    assert abs(bench.w_ref_tflop(bench.CONFIGS["sd15_c3lier"], 25) - 194.4) < 0.5
    assert abs(bench.w_ref_tflop(bench.CONFIGS["sdxl"], 25) - 816) < 2


def test_cpu_arm_projection_scales_only_the_denoise_loop():
    """the reference arm measures COMPLETE iterations at a small k; the workload's k is reached by scaling the measured
    per-denoise-step time, everything else (4 predictions, backward, AdamW) is taken as measured"""
    pr = bench.cpu_projection(bench.CONFIGS["sd21"], t_den=[12.0, 14.0], t_tot=[100.0, 104.0], k_meas=1, batch_meas=2, k=25)
    assert pr["measured"]["s_per_iteration"] == 102.0 and pr["measured"]["s_denoise_step"] == 13.0
    assert pr["measured"]["s_rest_of_iteration"] == 89.0
    assert abs(pr["projected"]["latents_per_s"] - 2 / (25 * 13.0 + 89.0)) < 1e-12
    assert abs(pr["measured"]["latents_per_s"] - 2 / 102.0) < 1e-12


REQUIRED = ["metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
            "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "roofline", "clocks"]


@pytest.mark.parametrize("name", ["r1_final_bench.json", "r1_bench_2gpu.json", "r2ai_final_bench.json"])
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


def test_reference_arm_runs_complete_iterations_and_keeps_the_contract():
    """`bench.py --impl reference` (reduced-width twin so it takes seconds here): the line is built from COMPLETE oracle
    iterations with measured phases; `steps` is the number actually timed, not the number requested."""
    import subprocess
    import sys
    pr = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", "tiny21",
                         "--steps", "20", "--warmup", "5"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert pr.returncode == 0, pr.stderr[-2000:]
    line = json.loads(pr.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "leco_train_latents_per_sec" and line["unit"] == "latents/s"
    assert line["steps"] == 2 and line["steps_requested"] == 20
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["value"] == line["value"] and cb["cores"] >= 1
    m, p = cb["measured"], cb["projected"]
    assert m["k"] == 1 and m["iterations"] == 2 and m["s_per_iteration"] > m["s_denoise_step"] > 0
    assert abs(p["latents_per_s"] - m["prompt_batch"] / (25 * m["s_denoise_step"] + m["s_rest_of_iteration"])) < 1e-9
    assert line["e2e"] == {"value": line["value"], "unit": "latents/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_both_arms_default_to_the_same_workload_and_idle_ranks_exit_clean():
    """N=1 -> BASELINE configs[1], N>1 -> configs[4] for our arm AND the reference arm; under torchrun only rank 0 of the
    reference arm works, the other ranks exit 0 without output; an SDXL reference request answers `unavailable`."""
    import subprocess
    import sys
    assert bench.default_workload(1) == "sd21" and bench.default_workload(2) == bench.default_workload(8) == "sd21_b4"
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    pr = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                        capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert pr.returncode == 0 and pr.stdout.strip() == ""
    pr = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", "sdxl"],
                        capture_output=True, text=True, timeout=120, cwd=ROOT)
    line = json.loads(pr.stdout.strip().splitlines()[-1])
    assert pr.returncode == 0 and line["impl"] == "reference" and "unavailable" in line


def test_committed_scaling_base_is_the_sharded_workload_on_one_gpu():
    """The N = 1 line (configs[1], prompt batch 2) and the N > 1 lines (configs[4] share, batch 4 per GPU) are different
    workloads; `scaling_base` in the N = 1 line is the latter on ONE GPU, so the 1 -> N curve has a like-for-like base."""
    import bench
    one = json.load(open(os.path.join(ROOT, "profiles", "r2am_bench_scaling_base.json")))
    two = json.load(open(os.path.join(ROOT, "profiles", "r2aj_bench_2gpu_sd21_b4_final.json")))
    sb = one["scaling_base"]
    assert one["config"]["name"] == bench.default_workload(1) and sb["name"] == bench.default_workload(2) == two["config"]["name"]
    assert sb["n_gpus"] == 1 and sb["unit"] == one["unit"]
    assert abs(sb["value"] - bench.CONFIGS[sb["name"]]["batch"] * 1000.0 / sb["ms_per_step"]) < 1e-6
    eff = two["value"] / (two["n_gpus"] * sb["value"])
    assert 0.9 < eff <= 1.02, eff          # like for like; value(2) / (2 x value(1)) of the mixed workloads would read 1.28
