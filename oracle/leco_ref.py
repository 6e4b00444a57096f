"""ORACLE (test infrastructure, never on the product path).

CPU restatement of LECO's OWN arithmetic for the training-step hot path, for use where
/root/reference does not exist (the GPU box).  Every function cites the reference lines
it follows.  PINNED: tests/golden/make_golden.py runs the reference's unmodified
train_lora.train() / lora.py / train_util.py / prompt_util.py in the build container and
tests/test_oracle_pinned.py checks this file against those outputs bit-for-bit (fp32,
same torch build), and against the committed fixtures everywhere else.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import List, Optional, Sequence

import torch
import torch.nn as nn

LORA_PREFIX = "lora_unet"                                   # lora.py:24
ATTN_TARGETS = ["Transformer2DModel"]                       # lora.py:15-17
CONV_TARGETS = ["ResnetBlock2D", "Downsample2D", "Upsample2D"]  # lora.py:18-22
METHOD_SKIP = {  # lora.py:170-187 — filters test the OUTER block name
    "noxattn": lambda n: "attn2" in n or "time_embed" in n,
    "innoxattn": lambda n: "attn2" in n,
    "selfattn": lambda n: "attn1" not in n,
    "xattn": lambda n: "attn2" not in n,
    "full": lambda n: False,
}


class LoRAModuleRef(nn.Module):
    """lora.py:44-106."""

    def __init__(self, lora_name: str, org: nn.Module, multiplier=1.0, lora_dim=4, alpha=1):
        super().__init__()
        self.lora_name = lora_name
        self.lora_dim = lora_dim
        kind = type(org).__name__
        if kind == "Linear":                                            # lora.py:62-66
            self.lora_down = nn.Linear(org.in_features, lora_dim, bias=False)
            self.lora_up = nn.Linear(lora_dim, org.out_features, bias=False)
        elif kind == "Conv2d":                                          # lora.py:68-82
            self.lora_dim = min(lora_dim, org.in_channels, org.out_channels)
            self.lora_down = nn.Conv2d(org.in_channels, self.lora_dim, org.kernel_size,
                                       org.stride, org.padding, bias=False)
            self.lora_up = nn.Conv2d(self.lora_dim, org.out_channels, (1, 1), (1, 1), bias=False)
        if isinstance(alpha, torch.Tensor):
            alpha = alpha.detach().numpy()
        alpha = lora_dim if (alpha is None or alpha == 0) else alpha    # lora.py:86
        self.scale = alpha / self.lora_dim                              # lora.py:87
        self.register_buffer("alpha", torch.tensor(alpha))              # lora.py:88
        nn.init.kaiming_uniform_(self.lora_down.weight, a=math.sqrt(5))  # lora.py:91
        nn.init.zeros_(self.lora_up.weight)                             # lora.py:92
        self.multiplier = multiplier
        self._org = [org]  # not registered as a submodule

    def apply_to(self):                                                 # lora.py:97-100
        org = self._org.pop()
        self.org_forward = org.forward
        org.forward = self.forward

    def forward(self, x):                                               # lora.py:102-106
        return self.org_forward(x) + self.lora_up(self.lora_down(x)) * self.multiplier * self.scale


class LoRANetworkRef(nn.Module):
    """lora.py:109-237.  `targets` stands for the module-level list the reference mutates in
    place for c3lier (train_lora.py:44-46)."""

    def __init__(self, unet: nn.Module, rank=4, multiplier=1.0, alpha=1.0, train_method="full",
                 targets: Optional[Sequence[str]] = None):
        super().__init__()
        self.multiplier, self.lora_dim, self.alpha = multiplier, rank, alpha
        targets = list(ATTN_TARGETS if targets is None else targets)
        if train_method not in METHOD_SKIP:
            raise NotImplementedError(f"train_method: {train_method} is not implemented.")
        skip = METHOD_SKIP[train_method]
        self.unet_loras: List[LoRAModuleRef] = []
        for name, module in unet.named_modules():                       # lora.py:169
            if skip(name) or type(module).__name__ not in targets:
                continue
            for child_name, child in module.named_modules():            # lora.py:189-197
                if type(child).__name__ in ("Linear", "Conv2d"):
                    lname = (LORA_PREFIX + "." + name + "." + child_name).replace(".", "_")
                    self.unet_loras.append(LoRAModuleRef(lname, child, multiplier, rank, alpha))
        names = [l.lora_name for l in self.unet_loras]
        assert len(set(names)) == len(names), "duplicated lora name"     # lora.py:139-144
        for lora in self.unet_loras:                                    # lora.py:147-152
            lora.apply_to()
            self.add_module(lora.lora_name, lora)

    def prepare_optimizer_params(self):                                 # lora.py:201-210
        if not self.unet_loras:
            return []
        return [{"params": [p for l in self.unet_loras for p in l.parameters()]}]

    def lora_state_dict(self, dtype=None):                              # lora.py:212-224
        sd = self.state_dict()
        out = {}
        for k, v in sd.items():
            if k.startswith("lora"):
                out[k] = v.detach().clone().to("cpu").to(dtype) if dtype is not None else v
        return out

    def __enter__(self):                                                # lora.py:231-233
        for l in self.unet_loras:
            l.multiplier = 1.0

    def __exit__(self, *exc):                                           # lora.py:235-237
        for l in self.unet_loras:
            l.multiplier = 0


# --------------------------------------------------------------------------- train_util
def get_initial_latents(scheduler, n_imgs, height, width, n_prompts, generator=None):
    """train_util.py:20-57: CPU randn(n,4,h/8,w/8), repeated per prompt, * init_noise_sigma."""
    noise = torch.randn((n_imgs, 4, height // 8, width // 8), generator=generator, device="cpu")
    return noise.repeat(n_prompts, 1, 1, 1) * scheduler.init_noise_sigma


def concat_embeddings(unconditional, conditional, n_imgs):
    """train_util.py:133-138: [u x n, c x n]."""
    return torch.cat([unconditional, conditional]).repeat_interleave(n_imgs, dim=0)


def predict_noise(unet, scheduler, timestep, latents, text_embeddings, guidance_scale=7.5,
                  added_cond_kwargs=None):
    """train_util.py:142-168 (and :217-257 for XL: the guidance rescale there is dead code,
    SURVEY Q6, so the XL twin returns the same guided value)."""
    x2 = scheduler.scale_model_input(torch.cat([latents] * 2), timestep)
    kw = {} if added_cond_kwargs is None else {"added_cond_kwargs": added_cond_kwargs}
    eps = unet(x2, timestep, encoder_hidden_states=text_embeddings, **kw).sample
    eps_u, eps_c = eps.chunk(2)
    return eps_u + guidance_scale * (eps_c - eps_u)


@torch.no_grad()
def diffusion(unet, scheduler, latents, text_embeddings, total_timesteps=1000, start_timesteps=0,
              **kw):
    """train_util.py:172-193."""
    for t in scheduler.timesteps[start_timesteps:total_timesteps]:
        eps = predict_noise(unet, scheduler, t, latents, text_embeddings, **kw)
        latents = scheduler.step(eps, t, latents).prev_sample
    return latents


def get_add_time_ids(height, width, dynamic_crops=False, dtype=torch.float32):
    """train_util.py:295-330 (static-crop branch; dynamic crops draw from the global RNG)."""
    if dynamic_crops:
        scale = torch.rand(1).item() * 2 + 1
        orig = (int(height * scale), int(width * scale))
        crop = (torch.randint(0, orig[0] - height, (1,)).item(),
                torch.randint(0, orig[1] - width, (1,)).item())
    else:
        orig, crop = (height, width), (0, 0)
    ids = list(orig + crop + (height, width))
    if 256 * len(ids) + 1280 != 2816:
        raise ValueError("add_time_ids length mismatch")
    return torch.tensor([ids], dtype=dtype)


def get_random_resolution_in_bucket(bucket_resolution=512):
    """train_util.py:404-416 (upper bound exclusive: never returns the bucket size)."""
    lo, hi = (bucket_resolution // 2) // 64, bucket_resolution // 64
    h = torch.randint(lo, hi, (1,)).item() * 64
    w = torch.randint(lo, hi, (1,)).item() * 64
    return h, w


# --------------------------------------------------------------------------- prompt_util
@dataclass
class PromptPairRef:
    """prompt_util.py:70-148 with the defaults of PromptSettings (:43-67)."""
    target: torch.Tensor
    positive: torch.Tensor
    unconditional: torch.Tensor
    neutral: torch.Tensor
    guidance_scale: float = 1.0
    resolution: int = 512
    dynamic_resolution: bool = False
    batch_size: int = 1
    action: str = "erase"
    dynamic_crops: bool = False      # XL only (prompt_util.py:53)

    def loss(self, target_latents, positive_latents, neutral_latents, unconditional_latents):
        if self.action == "erase":       # prompt_util.py:107-120
            goal = neutral_latents - self.guidance_scale * (positive_latents - unconditional_latents)
        elif self.action == "enhance":   # prompt_util.py:122-135
            goal = neutral_latents + self.guidance_scale * (positive_latents - unconditional_latents)
        else:
            raise ValueError("action must be erase or enhance")
        return torch.nn.functional.mse_loss(target_latents, goal)  # criteria = MSELoss, train_lora.py:96


# --------------------------------------------------------------------------- train_lora
def leco_iteration(unet, scheduler, network, optimizer, lr_scheduler, prompt_pairs, *,
                   max_denoising_steps=50, device="cpu", weight_dtype=torch.float32,
                   fixed_k: Optional[int] = None, record: Optional[dict] = None):
    """One pass of the loop body train_lora.py:141-302 (saving and logging excluded).
    RNG draw order is part of the contract: pair index, timesteps_to,
    [bucket h, w], latent noise — all from the global CPU generator.
    `record["clock"]` (a perf_counter-like callable, bench.py's reference arm) adds the wall time of the
    denoise loop (`t_denoise`) and of the whole iteration (`t_total`) to the record."""
    clk = record.get("clock") if record is not None else None
    t_begin = clk() if clk else 0.0
    with torch.no_grad():
        scheduler.set_timesteps(max_denoising_steps, device=device)          # :143-145
        optimizer.zero_grad()                                                # :147
        pair = prompt_pairs[torch.randint(0, len(prompt_pairs), (1,)).item()]  # :149-151
        k = torch.randint(1, max_denoising_steps, (1,)).item()               # :154-156
        if fixed_k is not None:
            k = fixed_k  # measurement variant only; the draw above still happens
        height = width = pair.resolution
        if pair.dynamic_resolution:                                          # :162-165
            height, width = get_random_resolution_in_bucket(pair.resolution)
        latents = get_initial_latents(scheduler, pair.batch_size, height, width, 1).to(
            device, dtype=weight_dtype)                                      # :175-177
        with network:                                                        # :179-193
            denoised = diffusion(unet, scheduler, latents,
                                 concat_embeddings(pair.unconditional, pair.target, pair.batch_size),
                                 start_timesteps=0, total_timesteps=k, guidance_scale=3)
        if clk:
            record["t_denoise"] = clk() - t_begin
        scheduler.set_timesteps(1000)                                        # :195
        t_cur = scheduler.timesteps[int(k * 1000 / max_denoising_steps)]     # :197-199

        def nograd_pass(cond):                                               # :202-237
            return predict_noise(unet, scheduler, t_cur, denoised,
                                 concat_embeddings(pair.unconditional, cond, pair.batch_size),
                                 guidance_scale=1).to("cpu", dtype=torch.float32)
        positive = nograd_pass(pair.positive)
        neutral = nograd_pass(pair.neutral)
        uncond = nograd_pass(pair.unconditional)
    with network:                                                            # :244-256
        target = predict_noise(unet, scheduler, t_cur, denoised,
                               concat_embeddings(pair.unconditional, pair.target, pair.batch_size),
                               guidance_scale=1).to("cpu", dtype=torch.float32)
    loss = pair.loss(target_latents=target, positive_latents=positive,
                     neutral_latents=neutral, unconditional_latents=uncond)   # :265-270
    loss.backward()                                                          # :279
    if record is not None:
        record.update(k=k, timestep=int(t_cur), height=height, width=width, denoised=denoised.detach().float().cpu(),
                      positive=positive, neutral=neutral, unconditional=uncond,
                      target=target.detach(), loss=float(loss.item()))
        if record.get("want_grads"):  # [d lora_down, d lora_up] per adapter, before the optimizer consumes them
            record["grads"] = [p.grad.detach().float().cpu().clone() for l in network.unet_loras
                               for p in (l.lora_down.weight, l.lora_up.weight)]
    optimizer.step()                                                         # :280
    lr_scheduler.step()                                                      # :281
    if clk:
        record["t_total"] = clk() - t_begin
    return float(loss.item())


class EmbedsXL:
    """prompt_util.PromptEmbedsXL (prompt_util.py:17-23): (text_embeds [1,77,2048], pooled_embeds [1,1280])."""

    def __init__(self, text_embeds, pooled_embeds):
        self.text_embeds, self.pooled_embeds = text_embeds, pooled_embeds


def leco_iteration_xl(unet, scheduler, network, optimizer, lr_scheduler, prompt_pairs, *, max_denoising_steps=50,
                      device="cpu", weight_dtype=torch.float32, fixed_k: Optional[int] = None,
                      record: Optional[dict] = None):
    """One pass of the SDXL loop body train_lora_xl.py:160-366: as leco_iteration plus the pooled text embedding
    and `add_time_ids` conditioning (train_util.py:217-291, :295-330).  PromptPairRef fields hold EmbedsXL."""
    with torch.no_grad():
        scheduler.set_timesteps(max_denoising_steps, device=device)           # :162-164
        optimizer.zero_grad()
        pair = prompt_pairs[torch.randint(0, len(prompt_pairs), (1,)).item()]   # :168-170
        k = torch.randint(1, max_denoising_steps, (1,)).item()                # :173-175
        if fixed_k is not None:
            k = fixed_k
        height = width = pair.resolution
        if pair.dynamic_resolution:                                           # :178-181
            height, width = get_random_resolution_in_bucket(pair.resolution)
        latents = get_initial_latents(scheduler, pair.batch_size, height, width, 1).to(device, dtype=weight_dtype)
        ids = get_add_time_ids(height, width, dynamic_crops=pair.dynamic_crops, dtype=weight_dtype).to(device)  # :196-201
        ids2 = concat_embeddings(ids, ids, pair.batch_size)

        def cond(c):
            return dict(text_embeddings=concat_embeddings(pair.unconditional.text_embeds, c.text_embeds, pair.batch_size),
                        added_cond_kwargs={"text_embeds": concat_embeddings(pair.unconditional.pooled_embeds,
                                                                            c.pooled_embeds, pair.batch_size),
                                           "time_ids": ids2})
        with network:                                                         # :203-225 (diffusion_xl)
            denoised = latents
            for t in scheduler.timesteps[0:k]:
                eps = predict_noise(unet, scheduler, t, denoised, guidance_scale=3, **cond(pair.target))
                denoised = scheduler.step(eps, t, denoised).prev_sample
        scheduler.set_timesteps(1000)
        t_cur = scheduler.timesteps[int(k * 1000 / max_denoising_steps)]

        def nograd_pass(c):
            return predict_noise(unet, scheduler, t_cur, denoised, guidance_scale=1, **cond(c)).to("cpu", dtype=torch.float32)
        positive, neutral, uncond = nograd_pass(pair.positive), nograd_pass(pair.neutral), nograd_pass(pair.unconditional)
    with network:
        target = predict_noise(unet, scheduler, t_cur, denoised, guidance_scale=1, **cond(pair.target)).to(
            "cpu", dtype=torch.float32)
    loss = pair.loss(target_latents=target, positive_latents=positive, neutral_latents=neutral,
                     unconditional_latents=uncond)
    loss.backward()
    if record is not None:
        record.update(k=k, timestep=int(t_cur), loss=float(loss.item()), height=height, width=width,
                      time_ids=[float(v) for v in ids.flatten().tolist()])
    optimizer.step()
    lr_scheduler.step()
    return float(loss.item())
