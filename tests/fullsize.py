"""Parity at the BASELINE size (test infrastructure).

BASELINE configs[1]: SD2.1 architecture, lierla rank 4, 64x64 latents (512 px), prompt batch 2.  Everything here
is a deterministic function of seeds, so the fp32 CPU oracle side is computed ONCE in the build container
(`python tests/fullsize.py`, ~30 min on 8 cores) and committed under tests/golden/cache/; the GPU tests load it.

Three actors run the SAME LECO iteration (same weights, same adapter values, same RNG draws, fixed k):
  * oracle_iteration_full()      fp32 CPU: oracle/leco_ref.leco_iteration on oracle/unet_ref (the pinned restatement of
                                 train_lora.py:141-302) — the reference answer;
  * torch_bf16_iteration_full()  the same oracle code moved to cuda in bf16 = the reference's own numerics
                                 (train_lora.py:67-78 runs the UNet and the adapters in bf16 through stock torch kernels);
  * engine_iteration_full()      the product: leco_b200.trainer.LecoTrainer on the CUDA kernels.
The tests bound the engine's distance from the fp32 answer by the distance of the plain-bf16 torch run (a MEASURED
tolerance, not a hand-set one).
"""
from __future__ import annotations

import contextlib
import io
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

ARCH, RES, BATCH, K_FIXED, MAX_STEPS, LR = "sd21", 512, 2, 3, 50, 1e-4
SEED_NET, SEED_UP, SEED_LOOP = 1234, 5, 7


def _embeddings(arch=ARCH):
    from leco_b200.synthetic import prompt_embedding
    from oracle.unet_ref import CONFIGS
    D = CONFIGS[arch].cross_attention_dim
    return {p: prompt_embedding(p, D).bfloat16().float() for p in ("van gogh", "")}


def seed_adapters(net):
    """bf16-representable adapter values: kaiming lora_down as drawn, NON-ZERO lora_up (0.05 N(0,1), seeded) so that
    d(lora_down) is non-zero and the LoRA-on denoise loop differs from the LoRA-off one.  Values are copied IN PLACE
    (the product network's parameters are views of its flat operand buffer)."""
    g = torch.Generator().manual_seed(SEED_UP)
    with torch.no_grad():
        for l in net.unet_loras:
            up = (0.05 * torch.randn(l.lora_up.weight.shape, generator=g)).bfloat16()
            l.lora_up.weight.copy_(up.to(device=l.lora_up.weight.device, dtype=l.lora_up.weight.dtype))
            dn = l.lora_down.weight.detach().to("cpu").bfloat16()
            l.lora_down.weight.copy_(dn.to(device=l.lora_down.weight.device, dtype=l.lora_down.weight.dtype))


def _oracle_iteration(device, dtype, arch=ARCH, k=K_FIXED, batch=BATCH, res=RES):
    from oracle import leco_ref
    from oracle.sched_ref import create_noise_scheduler
    from oracle.unet_ref import build_unet
    unet = build_unet(arch, seed=0).to(torch.bfloat16)       # the engine's bf16-rounded weights
    unet = unet.to(device=device, dtype=dtype)
    emb = {k_: v.to(device=device, dtype=dtype) for k_, v in _embeddings(arch).items()}
    torch.manual_seed(SEED_NET)
    with contextlib.redirect_stdout(io.StringIO()):
        net = leco_ref.LoRANetworkRef(unet, rank=4, multiplier=1.0, alpha=1.0)
    seed_adapters(net)
    net.to(device=device, dtype=dtype)
    pair = leco_ref.PromptPairRef(target=emb["van gogh"], positive=emb["van gogh"], unconditional=emb[""],
                                  neutral=emb[""], guidance_scale=1.0, resolution=res, batch_size=batch, action="erase")
    opt = torch.optim.AdamW(net.prepare_optimizer_params(), lr=LR)
    lrs = torch.optim.lr_scheduler.ConstantLR(opt, factor=1)
    sched = create_noise_scheduler("ddim", "v_prediction")
    before = [p.detach().float().cpu().clone() for l in net.unet_loras for p in (l.lora_down.weight, l.lora_up.weight)]
    torch.manual_seed(SEED_LOOP)
    rec = {"want_grads": True}
    loss = leco_ref.leco_iteration(unet, sched, net, opt, lrs, [pair], max_denoising_steps=MAX_STEPS, device=device,
                                   weight_dtype=dtype, fixed_k=k, record=rec)
    after = [p.detach().float().cpu() for l in net.unet_loras for p in (l.lora_down.weight, l.lora_up.weight)]
    return {"loss": loss, "k": rec["k"], "timestep": rec["timestep"], "denoised": rec["denoised"].float().cpu(),
            "target": rec["target"].float().cpu(), "positive": rec["positive"].float().cpu(),
            "neutral": rec["neutral"].float().cpu(),
            "grads": [g.to(torch.bfloat16) for g in rec["grads"]],                      # bf16: small fixture
            "update": [(a - b).to(torch.bfloat16) for a, b in zip(after, before)]}


def oracle_iteration_full(**kw):
    torch.set_num_threads(os.cpu_count() or 8)
    return _oracle_iteration("cpu", torch.float32, **kw)


def torch_bf16_iteration_full(**kw):
    return _oracle_iteration("cuda", torch.bfloat16, **kw)


def engine_trainer_full(arch=ARCH, batch=BATCH, res=RES, state_fp32=False, use_graphs=True):
    from leco_b200.lora import LoRANetwork
    from leco_b200.scheduler import DDIMScheduler
    from leco_b200.trainer import LecoTrainer, PromptPair
    from leco_b200.unet import SPECS, EngineUNet
    from oracle.unet_ref import build_unet
    eng = EngineUNet(SPECS[arch])
    eng.load_state_dict(build_unet(arch, seed=0).state_dict())
    eng.requires_grad_(False)
    eng.to("cuda")
    emb = _embeddings(arch)
    torch.manual_seed(SEED_NET)
    with contextlib.redirect_stdout(io.StringIO()):
        net = LoRANetwork(eng, rank=4, multiplier=1.0, alpha=1.0)
    net.to("cuda", dtype=torch.bfloat16)
    seed_adapters(net)
    pair = PromptPair(target=emb["van gogh"], positive=emb["van gogh"], unconditional=emb[""], neutral=emb[""],
                      guidance_scale=1.0, resolution=res, batch_size=batch, action="erase")
    trainer = LecoTrainer(eng, net, DDIMScheduler("v_prediction"), [pair], lr=LR, max_denoising_steps=MAX_STEPS,
                          use_cuda_graphs=use_graphs, state_fp32=state_fp32)
    return trainer, net


def engine_iteration_full(k=K_FIXED, **kw):
    trainer, net = engine_trainer_full(**kw)
    before = [p.detach().float().cpu().clone() for l in net.unet_loras for p in (l.lora_down.weight, l.lora_up.weight)]
    torch.manual_seed(SEED_LOOP)
    loss = trainer.iteration(fixed_k=k, step_optimizer=False)
    grads = [g.float().cpu() for g in net.adapter_grads()]
    trainer.optimizer.step()
    torch.cuda.synchronize()
    after = [p.detach().float().cpu() for l in net.unet_loras for p in (l.lora_down.weight, l.lora_up.weight)]
    return {"loss": loss.item(), "k": trainer.last["k"], "timestep": int(trainer.last["timestep"]),
            "denoised": trainer.last["denoised"].float().cpu(), "target": trainer.last["target"].float().cpu(),
            "grads": grads, "update": [a - b for a, b in zip(after, before)]}


# ------------------------------------------------------------------ comparison helpers
def rel_rms(a, b):
    a, b = a.float(), b.float()
    return ((a - b).pow(2).mean().sqrt() / b.pow(2).mean().sqrt().clamp_min(1e-30)).item()


def grads_rel_l2(got, ref):
    num = sum((g.float() - r.float()).pow(2).sum().item() for g, r in zip(got, ref))
    den = sum(r.float().pow(2).sum().item() for r in ref)
    return (num / den) ** 0.5


def grads_cosines(got, ref):
    """per-tensor cosine; tensors whose reference gradient is (numerically) zero are skipped."""
    out = []
    for g, r in zip(got, ref):
        g, r = g.float().reshape(-1), r.float().reshape(-1)
        if r.norm().item() < 1e-20:
            continue
        out.append((torch.dot(g, r) / (g.norm() * r.norm()).clamp_min(1e-30)).item())
    return out


def main():
    from tests.oracle_cache import cached
    import time
    t0 = time.time()
    from tests.gpu_checks import kernel_cases as kc
    if "--grads" in sys.argv or len(sys.argv) == 1:
        cached("grads_sd21_full", lambda: kc.oracle_grads("sd21", 2, 64, None), write=True)
        print("grads_sd21_full", round(time.time() - t0), "s", flush=True)
    if "--iter" in sys.argv or len(sys.argv) == 1:
        cached("iter_sd21_full", oracle_iteration_full, write=True)
        print("iter_sd21_full", round(time.time() - t0), "s", flush=True)


if __name__ == "__main__":
    main()
