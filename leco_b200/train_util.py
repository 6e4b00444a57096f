"""Step primitives with the reference's names and argument meaning (train_util.py), running on
the engine.  `predict_noise` / `diffusion` are what a drop-in user calls; they accept the
reference's own scheduler/unet duck types, so `train_lora.py`'s loop body works unchanged
when handed an `EngineUNet` + `leco_b200.scheduler.DDIMScheduler`.
"""
from __future__ import annotations

import torch

UNET_IN_CHANNELS = 4
VAE_SCALE_FACTOR = 8


def get_random_noise(batch_size: int, height: int, width: int, generator: torch.Generator = None) -> torch.Tensor:
    """train_util.py:20-32: CPU noise from the (global) generator."""
    return torch.randn((batch_size, UNET_IN_CHANNELS, height // VAE_SCALE_FACTOR, width // VAE_SCALE_FACTOR),
                       generator=generator, device="cpu")


def apply_noise_offset(latents: torch.Tensor, noise_offset: float) -> torch.Tensor:
    """train_util.py:36-40 (offset noise: one draw per sample and channel on the latents' device)."""
    return latents + noise_offset * torch.randn((latents.shape[0], latents.shape[1], 1, 1), device=latents.device)


def get_initial_latents(scheduler, n_imgs: int, height: int, width: int, n_prompts: int, generator=None):
    """train_util.py:43-57."""
    noise = get_random_noise(n_imgs, height, width, generator=generator).repeat(n_prompts, 1, 1, 1)
    return noise * scheduler.init_noise_sigma


def concat_embeddings(unconditional: torch.Tensor, conditional: torch.Tensor, n_imgs: int):
    """train_util.py:133-138: [uncond x n, cond x n]."""
    return torch.cat([unconditional, conditional]).repeat_interleave(n_imgs, dim=0)


def predict_noise(unet, scheduler, timestep, latents, text_embeddings, guidance_scale=7.5):
    """train_util.py:142-168: CFG-batched UNet call + guidance combine."""
    x2 = scheduler.scale_model_input(torch.cat([latents] * 2), timestep)
    eps = unet(x2, timestep, encoder_hidden_states=text_embeddings).sample
    eps_u, eps_c = eps.chunk(2)
    return eps_u + guidance_scale * (eps_c - eps_u)


@torch.no_grad()
def diffusion(unet, scheduler, latents, text_embeddings, total_timesteps: int = 1000, start_timesteps=0, **kwargs):
    """train_util.py:172-193."""
    for timestep in scheduler.timesteps[start_timesteps:total_timesteps]:
        eps = predict_noise(unet, scheduler, timestep, latents, text_embeddings, **kwargs)
        latents = scheduler.step(eps, timestep, latents).prev_sample
    return latents


def rescale_noise_cfg(noise_cfg: torch.Tensor, noise_pred_text: torch.Tensor, guidance_rescale: float = 0.0):
    """train_util.py:196-214 ("Common Diffusion Noise Schedules and Sample Steps are Flawed", section 3.4): match the
    guided prediction's per-sample std to the text branch's, then blend by `guidance_rescale`.  Plain torch on whatever
    device the predictions live on; the reference's XL loop computes it and discards the result, so it is
    not on the training path."""
    dims = list(range(1, noise_pred_text.ndim))
    std_text = noise_pred_text.std(dim=dims, keepdim=True)
    std_cfg = noise_cfg.std(dim=list(range(1, noise_cfg.ndim)), keepdim=True)
    rescaled = noise_cfg * (std_text / std_cfg)
    return guidance_rescale * rescaled + (1 - guidance_rescale) * noise_cfg


def predict_noise_xl(unet, scheduler, timestep, latents, text_embeddings, add_text_embeddings, add_time_ids,
                     guidance_scale=7.5, guidance_rescale=0.7):
    """train_util.py:217-257.  The reference computes a guidance rescale and then returns the
    un-rescaled guided value: so does this."""
    x2 = scheduler.scale_model_input(torch.cat([latents] * 2), timestep)
    eps = unet(x2, timestep, encoder_hidden_states=text_embeddings,
               added_cond_kwargs={"text_embeds": add_text_embeddings, "time_ids": add_time_ids}).sample
    eps_u, eps_c = eps.chunk(2)
    return eps_u + guidance_scale * (eps_c - eps_u)


@torch.no_grad()
def diffusion_xl(unet, scheduler, latents, text_embeddings, add_text_embeddings, add_time_ids,
                 guidance_scale: float = 1.0, total_timesteps: int = 1000, start_timesteps=0):
    """train_util.py:260-291."""
    for timestep in scheduler.timesteps[start_timesteps:total_timesteps]:
        eps = predict_noise_xl(unet, scheduler, timestep, latents, text_embeddings, add_text_embeddings,
                               add_time_ids, guidance_scale=guidance_scale, guidance_rescale=0.7)
        latents = scheduler.step(eps, timestep, latents).prev_sample
    return latents


def get_add_time_ids(height: int, width: int, dynamic_crops: bool = False, dtype: torch.dtype = torch.float32):
    """train_util.py:295-330."""
    if dynamic_crops:
        scale = torch.rand(1).item() * 2 + 1
        original = (int(height * scale), int(width * scale))
        crop = (torch.randint(0, original[0] - height, (1,)).item(),
                torch.randint(0, original[1] - width, (1,)).item())
    else:
        original, crop = (height, width), (0, 0)
    ids = list(original + crop + (height, width))
    if 256 * len(ids) + 1280 != 2816:
        raise ValueError("Model expects an added time embedding vector of length 2816")
    return torch.tensor([ids], dtype=dtype)


# --------------------------------------------------------------------------- optimizer / LR factories
class _LrProxy:
    """Drives torch.optim.lr_scheduler classes for an optimizer that is not a torch.optim.Optimizer (the fused flat
    optimizer): a real one-parameter SGD carries the schedule, its lr is mirrored into the fused optimizer."""

    def __init__(self, target, scheduler_ctor):
        self.target = target
        self._opt = torch.optim.SGD([torch.nn.Parameter(torch.zeros(1))], lr=float(target.lr))
        self._sched = scheduler_ctor(self._opt)
        self.target.lr = self._opt.param_groups[0]["lr"]

    def step(self):
        self._opt.step()
        self._sched.step()
        self.target.lr = self._opt.param_groups[0]["lr"]

    def get_last_lr(self):
        return self._sched.get_last_lr()


def get_optimizer(name: str):
    """train_util.py:333-370.  Returns a constructor `f(params_or_flat, lr=..., **kwargs)`.  adamw / adam / lion are
    the fused flat-buffer kernels; the adaptive-step-size families (dadapt*, prodigy) and the bitsandbytes 8-bit
    variants need third-party packages that neither the reference image nor this one ships."""
    from .lora import OPTIMIZER_MODES, FlatOptimizer
    name = name.lower()
    if name.startswith("dadapt"):
        if name not in ("dadaptadam", "dadaptlion"):
            raise ValueError("DAdapt optimizer must be dadaptadam or dadaptlion")
        raise NotImplementedError("dadaptation is not installed (optional in the reference too)")
    if name.endswith("8bit"):
        if name not in ("adam8bit", "lion8bit"):
            raise ValueError("8bit optimizer must be adam8bit or lion8bit")
        raise NotImplementedError("bitsandbytes is not installed (the reference marks this path unverified)")
    if name in OPTIMIZER_MODES:
        return lambda flat, **kw: FlatOptimizer(flat, name, **kw)
    if name == "prodigy":
        raise NotImplementedError("prodigyopt is not installed (optional in the reference too)")
    raise ValueError("Optimizer must be adam, adamw, lion or Prodigy")


def get_lr_scheduler(name, optimizer, max_iterations, lr_min, **kwargs):
    """train_util.py:373-401: the same torch.optim.lr_scheduler classes with the same arguments.  NB the reference's
    "linear" branch passes `factor=0.5` to LinearLR, whose parameter is `start_factor`: as shipped it raises
    TypeError; the evident intent (start at half the rate) is implemented."""
    S = torch.optim.lr_scheduler
    if name == "cosine":
        ctor = lambda o: S.CosineAnnealingLR(o, T_max=max_iterations, eta_min=lr_min, **kwargs)        # noqa: E731
    elif name == "cosine_with_restarts":
        ctor = lambda o: S.CosineAnnealingWarmRestarts(o, T_0=max_iterations // 10, T_mult=2, eta_min=lr_min, **kwargs)  # noqa: E731
    elif name == "step":
        ctor = lambda o: S.StepLR(o, step_size=max_iterations // 100, gamma=0.999, **kwargs)          # noqa: E731
    elif name == "constant":
        ctor = lambda o: S.ConstantLR(o, factor=1, **kwargs)                                           # noqa: E731
    elif name == "linear":
        ctor = lambda o: S.LinearLR(o, start_factor=0.5, total_iters=max_iterations // 100, **kwargs)  # noqa: E731
    else:
        raise ValueError("Scheduler must be cosine, cosine_with_restarts, step, linear or constant")
    if isinstance(optimizer, torch.optim.Optimizer):
        return ctor(optimizer)
    return _LrProxy(optimizer, ctor)


def get_random_resolution_in_bucket(bucket_resolution: int = 512):
    """train_util.py:404-416 (randint's upper bound is exclusive: the bucket size itself is never drawn)."""
    lo, hi = (bucket_resolution // 2) // 64, bucket_resolution // 64
    height = torch.randint(lo, hi, (1,)).item() * 64
    width = torch.randint(lo, hi, (1,)).item() * 64
    return height, width


# ---- the prompt-encoding functions of train_util.py:60-130 live beside the loaders (model_util); same names here ----
def text_tokenize(tokenizer, prompts):
    from . import model_util
    return model_util.text_tokenize(tokenizer, prompts)


def text_encode(text_encoder, tokens):
    from . import model_util
    return model_util.text_encode(text_encoder, tokens)


def encode_prompts(tokenizer, text_encoder, prompts):
    from . import model_util
    return model_util.encode_prompts(tokenizer, text_encoder, prompts)


def text_encode_xl(text_encoder, tokens, num_images_per_prompt: int = 1):
    from . import model_util
    return model_util.text_encode_xl(text_encoder, tokens, num_images_per_prompt)


def encode_prompts_xl(tokenizers, text_encoders, prompts, num_images_per_prompt: int = 1):
    from . import model_util
    return model_util.encode_prompts_xl(tokenizers, text_encoders, prompts, num_images_per_prompt)
