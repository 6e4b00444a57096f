"""ORACLE (test infrastructure, never on the product path).

Restatement of the four diffusers-0.20.0 schedulers the reference constructs at
model_util.py:237-274 (``DDIMScheduler`` / ``DDPMScheduler`` with
beta_start=0.00085, beta_end=0.012, "scaled_linear", 1000 train steps,
clip_sample=False; ``LMSDiscreteScheduler`` / ``EulerAncestralDiscreteScheduler`` with
the same betas) and uses at train_util.py:55 (init_noise_sigma), :153
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
    def step(self, model_output, timestep, sample, generator=None, noise=None):
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
            if noise is None:
                noise = torch.randn(model_output.shape, generator=generator,
                                    device=model_output.device, dtype=model_output.dtype)
            var = torch.clamp(b_prev / b_t * cur_beta, min=1e-20)  # fixed_small
            prev = prev + (var ** 0.5) * noise
        return SimpleNamespace(prev_sample=prev, pred_original_sample=x0)


class _SigmaBase:
    """Common part of the k-diffusion style schedulers (scheduling_lms_discrete.py / scheduling_euler_ancestral_discrete.py,
    diffusers 0.20.0): sigma = sqrt((1-abar)/abar); `set_timesteps` with the default "linspace" spacing interpolates the
    sigma table at n equally spaced (fractional) timesteps; the UNet input is sample / sqrt(sigma^2 + 1);
    init_noise_sigma = max sigma of the training table."""
    order = 1

    def __init__(self, prediction_type: str = "epsilon", num_train_timesteps: int = 1000):
        assert prediction_type in ("epsilon", "v_prediction")
        self.config = SimpleNamespace(prediction_type=prediction_type, num_train_timesteps=num_train_timesteps)
        self.alphas_cumprod = scaled_linear_alphas_cumprod(n=num_train_timesteps)
        sig = np.array(((1 - self.alphas_cumprod) / self.alphas_cumprod) ** 0.5)
        self._train_sigmas = sig
        self.sigmas = torch.from_numpy(np.concatenate([sig[::-1], [0.0]]).astype(np.float32))
        self.init_noise_sigma = self.sigmas.max()
        self.num_inference_steps = None
        self.timesteps = torch.from_numpy(np.linspace(0, num_train_timesteps - 1, num_train_timesteps, dtype=float)[::-1].copy())

    def set_timesteps(self, num_inference_steps: int, device=None):
        self.num_inference_steps = num_inference_steps
        T = self.config.num_train_timesteps
        ts = np.linspace(0, T - 1, num_inference_steps, dtype=float)[::-1].copy()
        sig = np.interp(ts, np.arange(0, len(self._train_sigmas)), self._train_sigmas)
        self.sigmas = torch.from_numpy(np.concatenate([sig, [0.0]]).astype(np.float32)).to(device)
        self.timesteps = torch.from_numpy(ts).to(device)
        self._reset()

    def _reset(self):
        pass

    def _index(self, timestep):
        t = float(timestep)
        idx = (self.timesteps.double().cpu() == t).nonzero()
        return int(idx[0].item())

    def scale_model_input(self, sample, timestep):
        sigma = self.sigmas[self._index(timestep)]
        return sample / ((sigma ** 2 + 1) ** 0.5)

    def _x0(self, model_output, sample, sigma):
        if self.config.prediction_type == "epsilon":
            return sample - sigma * model_output
        return model_output * (-sigma / (sigma ** 2 + 1) ** 0.5) + (sample / (sigma ** 2 + 1))


class EulerAncestralDiscreteScheduler(_SigmaBase):
    def step(self, model_output, timestep, sample, generator=None, noise=None):
        i = self._index(timestep)
        sigma = self.sigmas[i]
        x0 = self._x0(model_output, sample, sigma)
        sigma_from, sigma_to = self.sigmas[i], self.sigmas[i + 1]
        sigma_up = (sigma_to ** 2 * (sigma_from ** 2 - sigma_to ** 2) / sigma_from ** 2) ** 0.5
        sigma_down = (sigma_to ** 2 - sigma_up ** 2) ** 0.5
        derivative = (sample - x0) / sigma
        prev = sample + derivative * (sigma_down - sigma)
        if noise is None:
            noise = torch.randn(model_output.shape, generator=generator, device=model_output.device, dtype=model_output.dtype)
        prev = prev + noise * sigma_up
        return SimpleNamespace(prev_sample=prev, pred_original_sample=x0)


class LMSDiscreteScheduler(_SigmaBase):
    def _reset(self):
        self.derivatives = []

    def get_lms_coefficient(self, order, t, current_order):
        from scipy import integrate

        def lms_derivative(tau):
            prod = 1.0
            for k in range(order):
                if current_order == k:
                    continue
                prod *= (tau - self.sigmas[t - k]) / (self.sigmas[t - current_order] - self.sigmas[t - k])
            return prod
        return integrate.quad(lms_derivative, self.sigmas[t], self.sigmas[t + 1], epsrel=1e-4)[0]

    def step(self, model_output, timestep, sample, order: int = 4):
        i = self._index(timestep)
        sigma = self.sigmas[i]
        x0 = self._x0(model_output, sample, sigma)
        derivative = (sample - x0) / sigma
        self.derivatives.append(derivative)
        if len(self.derivatives) > order:
            self.derivatives.pop(0)
        order = min(i + 1, order)
        coeffs = [self.get_lms_coefficient(order, i, o) for o in range(order)]
        prev = sample + sum(c * d for c, d in zip(coeffs, reversed(self.derivatives)))
        return SimpleNamespace(prev_sample=prev, pred_original_sample=x0)


def create_noise_scheduler(name: str = "ddim", prediction_type: str = "epsilon"):
    """Same dispatch as model_util.py:230-278."""
    name = name.lower().replace(" ", "_")
    if name == "ddim":
        return DDIMScheduler(prediction_type)
    if name == "ddpm":
        return DDPMScheduler(prediction_type)
    if name == "lms":
        return LMSDiscreteScheduler(prediction_type)
    if name == "euler_a":
        return EulerAncestralDiscreteScheduler(prediction_type)
    raise ValueError(f"Unknown scheduler name: {name}")
