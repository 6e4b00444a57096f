"""GPU (B200): parity tests proper.  Every case calls the CUDA path through the C ABI
(leco_b200.ops -> ctypes -> libleco_b200.so) and compares with a plain-PyTorch fp32 reference
(tests/torch_backend.py) or the oracle (oracle/, fp32 CPU).  Tolerances are written in each case:
bf16 storage => 2e-2 of the reference's max-abs for single kernels, 3e-2 end to end."""
import pytest

from tests.gpu_checks import gemm_cases, kernel_cases

pytestmark = pytest.mark.gpu
from __graft_entry__ import assert_losses_close  # noqa: E402


def _bf16_yardstick(name, fn):
    """The same oracle iterations executed the way the reference executes them (bf16, stock torch kernels on this
    GPU): the measured tolerance of the loss comparisons (see __graft_entry__.loss_tolerance)."""
    import torch
    return fn(device="cuda", dtype=torch.bfloat16)

GEMM = [(n, f, kw) for n, f, kw in gemm_cases.CASES if not n.startswith(("perf_", "perfauto_"))]
KERN = [(n, f, kw) for n, f, kw in kernel_cases.CASES if n != "engine_fwd_sd21_64" and not n.startswith("perf_")]


@pytest.mark.parametrize("name,fn,kw", GEMM, ids=[c[0] for c in GEMM])
def test_tcgen05_gemm(name, fn, kw):
    res = fn(**kw)
    assert res["ok"], res


@pytest.mark.parametrize("name,fn,kw", KERN, ids=[c[0] for c in KERN])
def test_kernels_and_engine(name, fn, kw):
    res = fn(**kw)
    bad = {k: v for k, v in res.get("parts", {}).items() if not v.get("ok")}
    assert res["ok"], bad or res


def test_engine_forward_full_size_sd21():
    """BASELINE config size: SD2.1 architecture, 64x64 latents, CFG batch of 2 samples, vs the fp32 oracle."""
    res = kernel_cases.case_engine_forward("sd21", n=2, hw=64)   # oracle output: tests/golden/cache (fp32 CPU, ~1 min)
    assert res["ok"] and res["rel_rms"] < 2e-2, res


def test_engine_forward_full_size_sd15():
    """BASELINE configs[2] architecture: SD1.5 (8 heads: head dims 40 / 80 / 160, conv proj_in/out, context 768),
    64x64 latents, 2 samples, vs the fp32 oracle (committed fixture)."""
    res = kernel_cases.case_engine_forward("sd15", n=2, hw=64)
    assert res["ok"] and res["rel_rms"] < 2e-2, res


def test_engine_forward_full_size_sdxl():
    """BASELINE configs[3] architecture: SDXL (transformer depths 1/2/10, text_time conditioning), 128x128 latents
    (1024 px), 2 samples, vs the fp32 oracle (committed fixture)."""
    res = kernel_cases.case_engine_forward("sdxl", n=2, hw=128)
    assert res["ok"] and res["rel_rms"] < 2e-2, res


def test_engine_not_worse_than_reference_numerics():
    """The reference runs the UNet in bf16 (train_lora.py:67).  Measure how far a plain bf16 PyTorch
    execution of the oracle network is from the fp32 oracle and require the engine to be at least as
    close (within 1.5x): the engine's fused fp32-accumulate epilogues may not lose accuracy."""
    import torch
    from tests.oracle_cache import cached
    arch = "tiny21"
    ref = cached("fwd_tiny21_2_32", lambda: kernel_cases.oracle_forward(arch, 2, 32))
    oracle, eng = kernel_cases._engine_pair(arch)
    x, ctx, _ = kernel_cases._inputs(arch, 2, 32)
    t = torch.tensor(481)
    with torch.no_grad():
        out = eng(x.cuda(), t, encoder_hidden_states=ctx.cuda().bfloat16()).sample.float().cpu()
        ref_bf16 = oracle.to("cuda", torch.bfloat16)(x.cuda().bfloat16(), t.cuda(), ctx.cuda().bfloat16()).sample.float().cpu()
    rms = lambda a: (a - ref).pow(2).mean().sqrt().item() / ref.pow(2).mean().sqrt().item()  # noqa: E731
    e_engine, e_bf16 = rms(out), rms(ref_bf16)
    assert e_engine < 1.5 * e_bf16 + 1e-3, (e_engine, e_bf16)


@pytest.mark.parametrize("graphs,state_fp32,net_fp32", [(False, True, False), (True, True, False), (True, False, False),
                                                        (True, True, True)],
                         ids=["eager", "cuda_graphs", "cuda_graphs_bf16_optimizer_state", "cuda_graphs_float32_network"])
def test_leco_iteration_matches_oracle(graphs, state_fp32, net_fp32):
    """Three full LECO iterations (denoise loop, 4 predictions, erase loss, backward, AdamW) on the GPU
    vs oracle/leco_ref.leco_iteration (fp32 CPU, pinned against the reference's train loop) on the same
    seeds.  Tolerance: measured (bf16 torch execution of the same oracle, __graft_entry__.loss_tolerance).
    The third variant runs the default / benchmarked optimizer (bf16 moments, the reference's rounding points); the
    fourth `train.precision: float32` (fp32 master adapters stepped by leco_optim_flat_master)."""
    import torch
    from __graft_entry__ import engine_trainer, oracle_iterations
    from tests.oracle_cache import cached
    ref = cached("iters_tiny21", lambda: oracle_iterations(3))
    yard = _bf16_yardstick("iters_tiny21", lambda **kw: oracle_iterations(3, **kw))
    trainer, net = engine_trainer(use_graphs=graphs, state_fp32=state_fp32, net_dtype=torch.float32 if net_fp32 else None)
    if net_fp32:
        assert net.flat.master is not None and net.unet_loras[0].lora_up.weight.dtype == torch.float32
    torch.manual_seed(7)
    got, ks = [], []
    for _ in range(3):
        got.append(trainer.iteration().item())
        ks.append(trainer.last["k"])
    assert ks == ref["k"] == yard["k"]
    assert_losses_close(got, ref["losses"], yard["losses"])
    # adapter weights after 3 steps: same direction, same size
    num = den = 0.0
    for a, wb in zip(net.unet_loras, ref["lora_up"]):
        wa, wb = a.lora_up.weight.detach().float().cpu().reshape(-1), wb.float().reshape(-1)
        num += torch.dot(wa, wb).item()
        den += (wa.norm() * wb.norm()).item()
    assert num / den > 0.9, num / den
    if net_fp32:     # the operands the kernels read are the bf16 copy of the fp32 master, kept current by the optimizer
        assert torch.equal(net.flat.params, net.flat.master.bfloat16())


@pytest.mark.parametrize("graphs", [False, True], ids=["eager", "cuda_graphs"])
def test_leco_iteration_xl_matches_oracle(graphs):
    """three SDXL-loop iterations (pooled text embedding + add_time_ids with dynamic crops drawn
    after the noise, train_lora_xl.py:160-366) on the tinyxl topology vs oracle/leco_ref.leco_iteration_xl (pinned
    against the reference's train_lora_xl.train()).  Same k draws; loss within 5% (bf16 engine vs fp32 oracle)."""
    import torch
    from __graft_entry__ import engine_trainer_xl, oracle_iterations_xl
    from tests.oracle_cache import cached
    ref = cached("iters_tinyxl", lambda: oracle_iterations_xl(3))
    yard = _bf16_yardstick("iters_tinyxl", lambda **kw: oracle_iterations_xl(3, **kw))
    trainer, net = engine_trainer_xl(use_graphs=graphs)
    torch.manual_seed(7)
    got, ks = [], []
    for _ in range(3):
        got.append(trainer.iteration().item())
        ks.append(trainer.last["k"])
    assert ks == ref["k"]
    assert_losses_close(got, ref["losses"], yard["losses"])
    num = den = 0.0
    for a, wb in zip(net.unet_loras, ref["lora_up"]):
        wa, wb = a.lora_up.weight.detach().float().cpu().reshape(-1), wb.float().reshape(-1)
        num += torch.dot(wa, wb).item()
        den += (wa.norm() * wb.norm()).item()
    assert num / den > 0.9, num / den


def test_leco_iteration_lms_scheduler_matches_oracle():
    """train.noise_scheduler = "lms" (model_util.py:257-265): fractional timesteps, UNet input scaled by
    1/sqrt(sigma^2+1), init_noise_sigma = max sigma, 4-deep derivative history — through the fused trainer vs the oracle's
    restatement of LMSDiscreteScheduler driving oracle/leco_ref.leco_iteration."""
    import torch
    from __graft_entry__ import engine_trainer, oracle_iterations
    from tests.oracle_cache import cached
    ref = cached("iters_tiny21_lms", lambda: oracle_iterations(3, scheduler="lms"))
    yard = _bf16_yardstick("iters_tiny21_lms", lambda **kw: oracle_iterations(3, scheduler="lms", **kw))
    trainer, net = engine_trainer(use_graphs=True, scheduler="lms")
    torch.manual_seed(7)
    got, ks = [], []
    for _ in range(3):
        got.append(trainer.iteration().item())
        ks.append(trainer.last["k"])
    assert ks == ref["k"]
    assert_losses_close(got, ref["losses"], yard["losses"])


@pytest.mark.parametrize("name", ["ddpm", "euler_a"])
def test_leco_iteration_stochastic_schedulers_run(name):
    """DDPM / Euler-ancestral draw step noise on the device (diffusers: randn_tensor on the model's device), so an
    end-to-end comparison with a CPU oracle is not defined; their update rows are pinned on the CPU and the kernel
    against the rows (kernel_cases.case_sched_step).  Here: the fused trainer runs them, graphs replay, loss finite,
    and two identically seeded runs agree (device RNG included)."""
    import torch
    from __graft_entry__ import engine_trainer
    from leco_b200 import ops
    vals = []
    ops.set_deterministic(True)      # ordered reductions: with the device RNG seeded alike the two runs must be bit-equal
    try:
        for _ in range(2):
            trainer, net = engine_trainer(use_graphs=True, scheduler=name)
            torch.manual_seed(7)
            vals.append([trainer.iteration().item() for _ in range(2)])
    finally:
        ops.set_deterministic(False)
    assert all(v == v and v < 1e3 for v in vals[0]), vals
    assert vals[0] == vals[1], vals


def test_leco_iteration_dynamic_resolution_matches_oracle():
    """dynamic_resolution prompts (train_lora.py:162-165): the bucket draw picks a different, usually non-square latent
    size every iteration (here 24..40 on each side), so the trainer captures one graph pair per shape.  Same k draws,
    losses within 5 % + the bf16 floor of the oracle's."""
    import torch
    from __graft_entry__ import _SETTINGS_DYN, engine_trainer, oracle_iterations
    from tests.oracle_cache import cached
    ref = cached("iters_tiny21_dyn", lambda: oracle_iterations(4, settings=_SETTINGS_DYN))
    yard = _bf16_yardstick("iters_tiny21_dyn", lambda **kw: oracle_iterations(4, settings=_SETTINGS_DYN, **kw))
    trainer, net = engine_trainer(use_graphs=True, settings=_SETTINGS_DYN)
    torch.manual_seed(7)
    got, ks, shapes = [], [], []
    for _ in range(4):
        got.append(trainer.iteration().item())
        ks.append(trainer.last["k"])
        shapes.append(tuple(trainer.last["denoised"].shape[-2:]))
    assert ks == ref["k"]
    assert len(set(shapes)) > 1 and any(h != w for h, w in shapes), shapes
    assert_losses_close(got, ref["losses"], yard["losses"])


def test_leco_iteration_multiple_pairs_enhance_matches_oracle():
    """Two prompt pairs drawn at random per iteration (train_lora.py:149-151): an erase pair and an ENHANCE pair with
    guidance 1.5 and neutral != unconditional (three distinct LoRA-off prompts, prompt_util.py:119-135)."""
    import torch
    from __graft_entry__ import engine_trainer, oracle_iterations
    from tests.oracle_cache import cached
    ref = cached("iters_tiny21_multi", lambda: oracle_iterations(5, multi=True))
    yard = _bf16_yardstick("iters_tiny21_multi", lambda **kw: oracle_iterations(5, multi=True, **kw))
    trainer, net = engine_trainer(use_graphs=True, multi=True)
    torch.manual_seed(7)
    got, ks, acts = [], [], []
    for _ in range(5):
        got.append(trainer.iteration().item())
        ks.append(trainer.last["k"])
        acts.append(trainer.last["pair"].action)
    assert ks == ref["k"]
    assert "enhance" in acts and "erase" in acts, acts
    assert_losses_close(got, ref["losses"], yard["losses"])


# ----------------------------------------------------------------------------------------------------------------------
# Parity at the BASELINE size (configs[1]: SD2.1 architecture, 64x64 latents, prompt batch 2) — VERDICT r1 "next" #1.
# The fp32 oracle side is a committed fixture (tests/fullsize.py); the tolerance is MEASURED on this GPU by running the
# same oracle code in bf16 through stock torch kernels (= the reference's own numerics).
# ----------------------------------------------------------------------------------------------------------------------
def test_fullsize_sd21_lora_gradients():
    """LoRA gradients of one grad pass, SD2.1 full size, 2 samples at 64x64 (M = 8192 token rows at the top level):
    through the 2-CTA / split-K GEMMs, tn_reduce at full M and attention backward at S = 4096, vs fp32 oracle autograd."""
    res = kernel_cases.case_engine_grads("sd21", n=2, hw=64, cache_name="grads_sd21_full")
    g = res["parts"]["grads"]
    import json
    import os
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    if os.path.isdir(out):
        json.dump(res, open(os.path.join(out, "fullsize_grads_report.json"), "w"), indent=1)
    assert res["ok"] and g["rel"] < 5e-2 and g["cos_weighted"] > 0.99, res


def test_fullsize_sd21_iteration():
    """One complete LECO iteration at BASELINE configs[1] size with k = 3 denoise steps and non-zero adapters: denoised
    latents (row a7), target prediction, loss, every LoRA gradient, and the AdamW update (default bf16 optimizer state)
    against the fp32 oracle — each bounded by 1.5x the error of the plain-bf16 torch execution of the same oracle."""
    import torch
    from tests import fullsize as fs
    from tests.oracle_cache import cached
    ref = cached("iter_sd21_full", fs.oracle_iteration_full)
    yard = fs.torch_bf16_iteration_full()
    torch.cuda.empty_cache()
    got = fs.engine_iteration_full()
    assert got["k"] == ref["k"] == yard["k"] == fs.K_FIXED and got["timestep"] == ref["timestep"]
    report = {}
    for key in ("denoised", "target"):
        e, b = fs.rel_rms(got[key], ref[key]), fs.rel_rms(yard[key], ref[key])
        report[key] = (e, b)
        assert e <= 1.5 * b + 1e-3, (key, report)
    e, b = abs(got["loss"] - ref["loss"]) / ref["loss"], abs(yard["loss"] - ref["loss"]) / ref["loss"]
    report["loss"] = (got["loss"], yard["loss"], ref["loss"])
    assert e <= 1.5 * b + 0.02, report
    e, b = fs.grads_rel_l2(got["grads"], ref["grads"]), fs.grads_rel_l2(yard["grads"], ref["grads"])
    cos = fs.grads_cosines(got["grads"], ref["grads"])
    cos_b = fs.grads_cosines(yard["grads"], ref["grads"])
    report["grads"] = (e, b, min(cos), min(cos_b), sum(cos) / len(cos))
    assert e <= max(5e-2, 1.5 * b), report
    assert sum(cos) / len(cos) > 0.99 and min(cos) >= min(0.9, min(cos_b) - 0.05), report
    # AdamW update direction (bf16 parameters: an lr-sized step is ~1 ulp, so compare in aggregate)
    num = sum(torch.dot(a.reshape(-1).float(), b_.reshape(-1).float()).item() for a, b_ in zip(got["update"], yard["update"]))
    den = (sum(a.float().pow(2).sum().item() for a in got["update"]) * sum(b_.float().pow(2).sum().item() for b_ in yard["update"])) ** 0.5
    report["update_cos_vs_torch_bf16"] = num / den
    assert num / den > 0.9, report
    print("fullsize iteration report:", report)
    import json
    import os
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    if os.path.isdir(out):   # measured numbers for DESIGN.md / profiles (engine vs plain-bf16 torch, both against the fp32 oracle)
        json.dump({k: list(v) if isinstance(v, tuple) else v for k, v in report.items()},
                  open(os.path.join(out, "fullsize_iteration_report.json"), "w"), indent=1)


def test_load_weights_under_captured_graphs(tmp_path):
    """save_weights -> load_weights round trip on the GPU while CUDA graphs that captured the flat
    operand buffer are alive: values are copied in place, the next iteration uses them (resume)."""
    import torch
    from __graft_entry__ import engine_trainer
    from leco_b200 import ops
    ops.set_deterministic(True)      # ordered reductions: the two passes below must then agree to the bit
    try:
        _load_weights_round_trip(tmp_path, engine_trainer, torch)
    finally:
        ops.set_deterministic(False)


def _load_weights_round_trip(tmp_path, engine_trainer, torch):
    trainer, net = engine_trainer(use_graphs=True)
    torch.manual_seed(7)
    trainer.iteration()
    f = str(tmp_path / "ckpt.safetensors")
    net.save_weights(f, dtype=torch.bfloat16)
    snap = net.flat.params.clone()
    torch.manual_seed(99)
    rng = torch.get_rng_state()
    loss_a = trainer.iteration(step_optimizer=False).item()
    net.flat.grads.zero_()
    with torch.no_grad():
        net.flat.params.mul_(0.5)                      # wreck the adapters ...
    net.load_weights(f)                                 # ... and restore them from the file, in place
    assert torch.equal(net.flat.params, snap)
    torch.set_rng_state(rng)
    loss_b = trainer.iteration(step_optimizer=False).item()
    assert loss_a == loss_b, (loss_a, loss_b)      # same graphs, same restored weights, ordered reductions


def test_lr_schedule_reaches_the_fused_optimizer():
    """train_lora.py:281 steps the LR scheduler every iteration: cosine schedule through LecoTrainer."""
    import math
    import torch
    from __graft_entry__ import engine_trainer
    trainer, net = engine_trainer(use_graphs=True)
    from leco_b200.train_util import get_lr_scheduler
    trainer.lr_scheduler = get_lr_scheduler("cosine", trainer.optimizer, max_iterations=10, lr_min=1e-5)
    lr0 = trainer.optimizer.lr
    torch.manual_seed(7)
    lrs = []
    for _ in range(3):
        trainer.iteration()
        lrs.append(trainer.last["lr"])
    expect = [1e-5 + (lr0 - 1e-5) * (1 + math.cos(math.pi * t / 10)) / 2 for t in (1, 2, 3)]
    assert all(abs(a - b) < 1e-9 for a, b in zip(lrs, expect)), (lrs, expect)


def test_data_parallel_two_gpus_nccl():
    """SURVEY §8e on real hardware: 2 ranks x local batch 1 through LecoTrainer.iteration's NCCL branch (one all-reduce
    of gradient + loss) vs one rank at batch 2.  Needs 2 GPUs (skipped on the 1-GPU test box; `gpurun --gpus 2`)."""
    import json
    import os
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = os.path.join(root, "gpurun_out", "dp_check_w2.json")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29531", os.path.join(root, "tests", "gpu_checks", "dp_check.py")]
    pr = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=root)
    assert pr.returncode == 0, pr.stderr[-3000:]
    res = json.load(open(out))
    assert res["ok"], res


@pytest.mark.parametrize("precision", ["bfloat16", "float32"])
def test_train_driver_runs_an_examples_style_config(tmp_path, precision):
    """`leco_b200.train_lora.train(config, prompts)` = the reference's driver (train_lora.py:34-321) on the fused trainer:
    a YAML with exactly the keys of examples/config.yaml / examples/prompts.yaml (architecture swapped for the
    reduced-width twin, lion + cosine as in examples/unreal_config.yaml) trains, follows the cosine LR schedule, writes
    the periodic and final .safetensors files with the kohya key set, and the saved file loads back.  precision =
    float32 is the notebook's setting (train.ipynb): fp32 master adapters, saved as fp32 (train_lora.py:55)."""
    import torch
    import yaml
    from leco_b200 import config_util, train_lora
    from safetensors.torch import load_file
    cfg = {"prompts_file": str(tmp_path / "prompts.yaml"),
           "pretrained_model": {"name_or_path": "tiny21", "v2": True, "v_pred": True},
           "network": {"type": "lierla", "rank": 4, "alpha": 1.0, "training_method": "full"},
           "train": {"precision": precision, "noise_scheduler": "ddim", "iterations": 5, "lr": "1e-4", "optimizer": "lion",
                     "optimizer_args": "weight_decay=0.01", "lr_scheduler": "cosine", "max_denoising_steps": 8},
           "save": {"name": "van_gogh", "path": str(tmp_path / "output"), "per_steps": 2, "precision": "bfloat16"},
           "logging": {"use_wandb": False, "verbose": False}, "other": {"use_xformers": True}}
    prompts = [{"target": "van gogh", "positive": "van gogh", "unconditional": "", "neutral": "", "action": "erase",
                "guidance_scale": 1.0, "resolution": 128, "dynamic_resolution": False, "batch_size": 2}]
    (tmp_path / "config.yaml").write_text(yaml.safe_dump(cfg))
    (tmp_path / "prompts.yaml").write_text(yaml.safe_dump(prompts))
    config = config_util.load_config_from_yaml(str(tmp_path / "config.yaml"))
    settings = config_util.load_prompts_from_yaml(config.prompts_file)
    torch.manual_seed(7)
    seen = []
    losses = train_lora.train(config, settings, on_iteration=lambda i, v: seen.append((i, v)))
    assert len(losses) == 5 and all(v == v and 0 < v < 10 for v in losses) and [i for i, _ in seen] == list(range(5))
    out = sorted(p.name for p in (tmp_path / "output").iterdir())
    assert out == ["van_gogh_2steps.safetensors", "van_gogh_4steps.safetensors", "van_gogh_last.safetensors"] or \
        out == ["van_gogh_2steps.safetensors", "van_gogh_last.safetensors"], out     # i == iterations-1 is skipped
    sd = load_file(str(tmp_path / "output" / "van_gogh_last.safetensors"))
    want = config_util.parse_precision(precision)
    assert len(sd) == 192 * 3 and all(v.dtype == want for k, v in sd.items() if "lora_" in k.split(".")[-2])
    up = [v for k, v in sd.items() if k.endswith("lora_up.weight")]
    assert any(float(u.abs().max()) > 0 for u in up)           # lion moved the zero-initialised lora_up


@pytest.mark.parametrize("arch", ["tiny21", "tinyxl"])
def test_train_driver_from_a_checkpoint_directory(tmp_path, arch):
    """SURVEY §8f rank 1 end to end: a diffusers-layout checkpoint directory (synthetic weights: nothing real exists
    offline) -> `load_models[_xl]` reads topology, UNet weights, tokenizer(s) and CLIP text encoder(s) -> prompts are
    tokenized and encoded ON THE GPU by the engine's text encoder (train_util.py:60-130) -> the loop trains.  The
    embeddings the trainer sees are checked against the fp32 oracle on the same tokens."""
    import torch
    import yaml
    from leco_b200 import config_util, model_util, train_lora
    from leco_b200.unet import SPECS
    from oracle import clip_ref
    from tests.clip_fixtures import write_checkpoint_dir
    d = write_checkpoint_dir(str(tmp_path / "ckpt"), arch, seed=5)
    xl = SPECS[arch].text_time
    prompts = ["van gogh style painting!", ""]
    if xl:
        toks, encs, unet, _ = model_util.load_models_xl(d)
        out = model_util.encode_prompts_xl(toks, encs, prompts)
        text, pooled = out.text_embeds, out.pooled_embeds
    else:
        tok, enc, unet, _ = model_util.load_models(d, v2=True, v_pred=True)
        toks, encs = [tok], [enc]
        text, pooled = model_util.encode_prompts(tok, enc, prompts), None
    assert text.is_cuda and text.dtype == torch.bfloat16 and text.shape == (2, 77, SPECS[arch].cross_attention_dim)
    parts = []
    for t, e in zip(toks, encs):
        ids = model_util.text_tokenize(t, prompts)
        sd = {k: v.float().cpu() for k, v in e.state_dict().items()}
        last, _, emb, hidden = clip_ref.clip_text_forward(sd, ids, heads=e.spec.num_attention_heads, act=e.spec.hidden_act,
                                                          eos_token_id=e.spec.eos_token_id)
        parts.append(hidden[-2] if xl else last)
    want = torch.cat(parts, -1)
    rel = ((text.float().cpu() - want).pow(2).mean().sqrt() / want.pow(2).mean().sqrt()).item()
    assert rel < 1.5e-2, rel                                     # bf16 activations through 2-3 layers
    if xl:
        relp = ((pooled.float().cpu() - emb).pow(2).mean().sqrt() / emb.pow(2).mean().sqrt()).item()
        assert pooled.shape == (2, SPECS[arch].add_text_dim) and relp < 1.5e-2, relp
    del unet, encs

    cfg = {"prompts_file": str(tmp_path / "prompts.yaml"),
           "pretrained_model": {"name_or_path": d, "v2": not xl, "v_pred": not xl},
           "network": {"type": "lierla", "rank": 4, "alpha": 1.0, "training_method": "full"},
           "train": {"precision": "bfloat16", "noise_scheduler": "ddim", "iterations": 3, "lr": "1e-4", "optimizer": "adamw",
                     "lr_scheduler": "constant", "max_denoising_steps": 8},
           "save": {"name": "out", "path": str(tmp_path / "output"), "per_steps": 200, "precision": "bfloat16"},
           "logging": {"use_wandb": False, "verbose": False}, "other": {"use_xformers": True}}
    pr = [{"target": "van gogh", "positive": "van gogh", "unconditional": "", "neutral": "", "action": "erase",
           "guidance_scale": 1.0, "resolution": 128, "dynamic_resolution": False, "batch_size": 1}]
    (tmp_path / "config.yaml").write_text(yaml.safe_dump(cfg))
    (tmp_path / "prompts.yaml").write_text(yaml.safe_dump(pr))
    config = config_util.load_config_from_yaml(str(tmp_path / "config.yaml"))
    settings = config_util.load_prompts_from_yaml(config.prompts_file)
    torch.manual_seed(3)
    seen = []
    losses = train_lora.train(config, settings, xl=xl, on_iteration=lambda i, v: seen.append(v))
    assert len(losses) == 3 and all(v == v and 0 < v < 50 for v in losses), losses
    assert (tmp_path / "output" / "out_last.safetensors").is_file()


@pytest.mark.parametrize("net_fp32", [False, True], ids=["bfloat16", "float32_network"])
def test_reference_loop_body_drives_the_engine_through_the_drop_in_surface(net_fp32):
    """the reference's loop body — train_lora.py:141-302 as restated statement for statement by
    oracle/leco_ref.leco_iteration, which tests/test_oracle_pinned.py holds bit-equal to the reference's own train() —
    runs UNMODIFIED on the GPU with the engine behind the reference's call surface: `unet(...).sample` inside
    predict_noise / diffusion, `scheduler.step(...).prev_sample`, `with network:`, `loss.backward()`, torch.optim.AdamW
    on the adapters' Parameters (the reference's files themselves do not travel to the GPU box).
    Same k draws as the fp32 oracle, losses inside the measured bf16 tolerance, adapters move the same way.
    float32_network: `train.precision: float32` — torch's AdamW steps fp32 Parameters (views of the flat fp32 master),
    the engine re-derives its bf16 operands before every pass."""
    import torch
    from __graft_entry__ import dropin_iterations, oracle_iterations
    from tests.oracle_cache import cached
    ref = cached("iters_tiny21", lambda: oracle_iterations(3))
    yard = _bf16_yardstick("iters_tiny21", lambda **kw: oracle_iterations(3, **kw))
    got = dropin_iterations(3, net_dtype=torch.float32 if net_fp32 else None)
    assert got["k"] == ref["k"]
    assert_losses_close(got["losses"], ref["losses"], yard["losses"])
    num = den = 0.0
    for wa, wb in zip(got["lora_up"], ref["lora_up"]):
        wa, wb = wa.float().reshape(-1), wb.float().reshape(-1)
        num += torch.dot(wa, wb).item()
        den += (wa.norm() * wb.norm()).item()
    assert den > 0 and num / den > 0.9, num / den
