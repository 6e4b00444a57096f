"""ORACLE (test infrastructure, never on the product path).

Restatement of the two diffusers-0.20.0 schedulers the reference constructs at
model_util.py:237-256 (``DDIMScheduler`` / ``DDPMScheduler`` with
beta_start=0.00085, beta_end=0.012, "scaled_linear", 1000 train steps,
clip_sample=False) and uses at train_util.py:55 (init_noise_sigma), :153
(scale_model_input), :190 (step().prev_sample), train_lora.py:143-145,195-199
(set_timesteps / timesteps).  diffusers is absent from /root/reference and from
this image: PARITY UNPINNED for the third-party part; the closed forms below are
the published algorithm (schedulers/scheduling_ddim.py, scheduling_ddpm.py).
"""
from __future__ import annotations

from types import SimpleNamespace

import numpy as np
import torch


def scaled_linear_alphas_cumprod(beta_start=0.00085, beta_end=0.012, n=1000) -> torch.Tensor:
    betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, n, dtype=torch.float32) ** 2
    return torch.cumprod(1.0 - betas, dim=0)


class _Base:
    init_noise_sigma = 1.0
    order = 1

    def __init__(self, prediction_type: str = "epsilon", num_train_timesteps: int = 1000):
        assert prediction_type in ("epsilon", "v_prediction")
        self.config = SimpleNamespace(prediction_type=prediction_type,
                                      num_train_timesteps=num_train_timesteps)
        self.alphas_cumprod = scaled_linear_alphas_cumprod(n=num_train_timesteps)
        self.final_alpha_cumprod = torch.tensor(1.0)  # set_alpha_to_one=True
        self.num_inference_steps = None
        self.timesteps = torch.from_numpy(np.arange(0, num_train_timesteps)[::-1].copy())

    def set_timesteps(self, num_inference_steps: int, device=None):
        # timestep_spacing="leading", steps_offset=0
        self.num_inference_steps = num_inference_steps
        ratio = self.config.num_train_timesteps // num_inference_steps
        ts = (np.arange(0, num_inference_steps) * ratio).round()[::-1].copy().astype(np.int64)
        self.timesteps = torch.from_numpy(ts).to(device)

    def scale_model_input(self, sample, timestep=None):
        return sample

    def _x0_eps(self, model_output, sample, a_t):
        b_t = 1 - a_t
        if self.config.prediction_type == "epsilon":
            x0 = (sample - b_t ** 0.5 * model_output) / a_t ** 0.5
            eps = model_output
        else:
            x0 = (a_t ** 0.5) * sample - (b_t ** 0.5) * model_output
            eps = (a_t ** 0.5) * model_output + (b_t ** 0.5) * sample
        return x0, eps


class DDIMScheduler(_Base):
    def step(self, model_output, timestep, sample, eta: float = 0.0):
        t = int(timestep)
        prev_t = t - self.config.num_train_timesteps // self.num_inference_steps
        a_t = self.alphas_cumprod[t]
        a_prev = self.alphas_cumprod[prev_t] if prev_t >= 0 else self.final_alpha_cumprod
        x0, eps = self._x0_eps(model_output, sample, a_t)
        # eta = 0 -> std_dev_t = 0
        direction = (1 - a_prev) ** 0.5 * eps
        prev_sample = a_prev ** 0.5 * x0 + direction
        return SimpleNamespace(prev_sample=prev_sample, pred_original_sample=x0)


class DDPMScheduler(_Base):
    def step(self, model_output, timestep, sample, generator=None):
        t = int(timestep)
        n_inf = self.num_inference_steps or self.config.num_train_timesteps
        prev_t = t - self.config.num_train_timesteps // n_inf
        a_t = self.alphas_cumprod[t]
        a_prev = self.alphas_cumprod[prev_t] if prev_t >= 0 else torch.tensor(1.0)
        b_t, b_prev = 1 - a_t, 1 - a_prev
        cur_alpha = a_t / a_prev
        cur_beta = 1 - cur_alpha
        x0, _ = self._x0_eps(model_output, sample, a_t)
        x0_coeff = (a_prev ** 0.5 * cur_beta) / b_t
        x_coeff = cur_alpha ** 0.5 * b_prev / b_t
        prev = x0_coeff * x0 + x_coeff * sample
        if t > 0:
            noise = torch.randn(model_output.shape, generator=generator,
                                device=model_output.device, dtype=model_output.dtype)
            var = torch.clamp(b_prev / b_t * cur_beta, min=1e-20)  # fixed_small
            prev = prev + (var ** 0.5) * noise
        return SimpleNamespace(prev_sample=prev, pred_original_sample=x0)


def create_noise_scheduler(name: str = "ddim", prediction_type: str = "epsilon"):
    """Same dispatch as model_util.py:230-278 for the two schedulers in scope."""
    name = name.lower().replace(" ", "_")
    if name == "ddim":
        return DDIMScheduler(prediction_type)
    if name == "ddpm":
        return DDPMScheduler(prediction_type)
    raise ValueError(f"Unknown scheduler name: {name}")
