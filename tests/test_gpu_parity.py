"""GPU (B200): parity tests proper.  Every case calls the CUDA path through the C ABI
(leco_b200.ops -> ctypes -> libleco_b200.so) and compares with a plain-PyTorch fp32 reference
(tests/torch_backend.py) or the oracle (oracle/, fp32 CPU).  Tolerances are written in each case:
bf16 storage => 2e-2 of the reference's max-abs for single kernels, 3e-2 end to end."""
import pytest

from tests.gpu_checks import gemm_cases, kernel_cases

pytestmark = pytest.mark.gpu
from __graft_entry__ import BF16_LOSS_FLOOR  # noqa: E402

GEMM = [(n, f, kw) for n, f, kw in gemm_cases.CASES if not n.startswith(("perf_", "perfauto_"))]
KERN = [(n, f, kw) for n, f, kw in kernel_cases.CASES if n != "engine_fwd_sd21_64"]


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


@pytest.mark.parametrize("graphs", [False, True], ids=["eager", "cuda_graphs"])
def test_leco_iteration_matches_oracle(graphs):
    """Three full LECO iterations (denoise loop, 4 predictions, erase loss, backward, AdamW) on the GPU
    vs oracle/leco_ref.leco_iteration (fp32 CPU, pinned against the reference's train loop) on the same
    seeds.  Tolerance: 5% of the loss (bf16 network vs fp32 oracle)."""
    import torch
    from __graft_entry__ import engine_trainer, oracle_iterations
    from tests.oracle_cache import cached
    ref = cached("iters_tiny21", lambda: oracle_iterations(3))
    trainer, net = engine_trainer(use_graphs=graphs)
    torch.manual_seed(7)
    got, ks = [], []
    for _ in range(3):
        got.append(trainer.iteration().item())
        ks.append(trainer.last["k"])
    assert ks == ref["k"]
    for a, b in zip(got, ref["losses"]):
        assert abs(a - b) <= 0.05 * abs(b) + BF16_LOSS_FLOOR, (got, ref["losses"])
    # adapter weights after 3 steps: same direction, same size
    num = den = 0.0
    for a, wb in zip(net.unet_loras, ref["lora_up"]):
        wa, wb = a.lora_up.weight.detach().float().cpu().reshape(-1), wb.float().reshape(-1)
        num += torch.dot(wa, wb).item()
        den += (wa.norm() * wb.norm()).item()
    assert num / den > 0.9, num / den


@pytest.mark.parametrize("graphs", [False, True], ids=["eager", "cuda_graphs"])
def test_leco_iteration_xl_matches_oracle(graphs):
    """three SDXL-loop iterations (pooled text embedding + add_time_ids with dynamic crops drawn
    after the noise, train_lora_xl.py:160-366) on the tinyxl topology vs oracle/leco_ref.leco_iteration_xl (pinned
    against the reference's train_lora_xl.train()).  Same k draws; loss within 5% (bf16 engine vs fp32 oracle)."""
    import torch
    from __graft_entry__ import engine_trainer_xl, oracle_iterations_xl
    from tests.oracle_cache import cached
    ref = cached("iters_tinyxl", lambda: oracle_iterations_xl(3))
    trainer, net = engine_trainer_xl(use_graphs=graphs)
    torch.manual_seed(7)
    got, ks = [], []
    for _ in range(3):
        got.append(trainer.iteration().item())
        ks.append(trainer.last["k"])
    assert ks == ref["k"]
    for a, b in zip(got, ref["losses"]):
        assert abs(a - b) <= 0.05 * abs(b) + BF16_LOSS_FLOOR, (got, ref["losses"])
    num = den = 0.0
    for a, wb in zip(net.unet_loras, ref["lora_up"]):
        wa, wb = a.lora_up.weight.detach().float().cpu().reshape(-1), wb.float().reshape(-1)
        num += torch.dot(wa, wb).item()
        den += (wa.norm() * wb.norm()).item()
    assert num / den > 0.9, num / den


def test_leco_iteration_dynamic_resolution_matches_oracle():
    """dynamic_resolution prompts (train_lora.py:162-165): the bucket draw picks a different, usually non-square latent
    size every iteration (here 24..40 on each side), so the trainer captures one graph pair per shape.  Same k draws,
    losses within 5 % + the bf16 floor of the oracle's."""
    import torch
    from __graft_entry__ import _SETTINGS_DYN, engine_trainer, oracle_iterations
    from tests.oracle_cache import cached
    ref = cached("iters_tiny21_dyn", lambda: oracle_iterations(4, settings=_SETTINGS_DYN))
    trainer, net = engine_trainer(use_graphs=True, settings=_SETTINGS_DYN)
    torch.manual_seed(7)
    got, ks, shapes = [], [], []
    for _ in range(4):
        got.append(trainer.iteration().item())
        ks.append(trainer.last["k"])
        shapes.append(tuple(trainer.last["denoised"].shape[-2:]))
    assert ks == ref["k"]
    assert len(set(shapes)) > 1 and any(h != w for h, w in shapes), shapes
    for a, b in zip(got, ref["losses"]):
        assert abs(a - b) <= 0.05 * abs(b) + BF16_LOSS_FLOOR, (got, ref["losses"])


def test_leco_iteration_multiple_pairs_enhance_matches_oracle():
    """Two prompt pairs drawn at random per iteration (train_lora.py:149-151): an erase pair and an ENHANCE pair with
    guidance 1.5 and neutral != unconditional (three distinct LoRA-off prompts, prompt_util.py:119-135)."""
    import torch
    from __graft_entry__ import engine_trainer, oracle_iterations
    from tests.oracle_cache import cached
    ref = cached("iters_tiny21_multi", lambda: oracle_iterations(5, multi=True))
    trainer, net = engine_trainer(use_graphs=True, multi=True)
    torch.manual_seed(7)
    got, ks, acts = [], [], []
    for _ in range(5):
        got.append(trainer.iteration().item())
        ks.append(trainer.last["k"])
        acts.append(trainer.last["pair"].action)
    assert ks == ref["k"]
    assert "enhance" in acts and "erase" in acts, acts
    for a, b in zip(got, ref["losses"]):
        assert abs(a - b) <= 0.05 * abs(b) + BF16_LOSS_FLOOR, (got, ref["losses"])
