"""DDIM scheduler with the diffusers call surface the reference uses.

Mirrors what `model_util.create_noise_scheduler("ddim", prediction_type)` returns
(model_util.py:237-246: scaled-linear betas 0.00085..0.012, 1000 train steps,
clip_sample=False, set_alpha_to_one) for the calls on the hot path:
  set_timesteps / timesteps        train_lora.py:143-145, 195-199
  scale_model_input (identity)     train_util.py:153
  step(...).prev_sample  (eta=0)   train_util.py:190
  init_noise_sigma (= 1.0)         train_util.py:55
The update is an affine map x' = cx*x + ce*model_output for both prediction types, so the
fused trainer folds it into the CFG-combine kernel (ops.guided_step); `step()` here applies
the same coefficients for drop-in callers.
"""
from __future__ import annotations

import math
from types import SimpleNamespace
from typing import List, Tuple

import torch


def _alphas_cumprod(n=1000, beta_start=0.00085, beta_end=0.012) -> List[float]:
    # float32 table exactly like torch.linspace(...)**2 -> cumprod (diffusers builds it in fp32)
    betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, n, dtype=torch.float32) ** 2
    return torch.cumprod(1.0 - betas, dim=0).double().tolist()


class DDIMScheduler:
    init_noise_sigma = 1.0
    order = 1

    def __init__(self, prediction_type: str = "epsilon", num_train_timesteps: int = 1000):
        if prediction_type not in ("epsilon", "v_prediction"):
            raise ValueError(f"unsupported prediction_type {prediction_type}")
        self.config = SimpleNamespace(prediction_type=prediction_type, num_train_timesteps=num_train_timesteps)
        self._acp = _alphas_cumprod(num_train_timesteps)
        self.alphas_cumprod = torch.tensor(self._acp, dtype=torch.float32)
        self.final_alpha_cumprod = 1.0
        self.num_inference_steps = None
        self.timesteps = torch.arange(num_train_timesteps - 1, -1, -1, dtype=torch.int64)

    def set_timesteps(self, num_inference_steps: int, device=None):
        """'leading' spacing, steps_offset 0:  t_i = (n-1-i) * (T // n)."""
        self.num_inference_steps = int(num_inference_steps)
        ratio = self.config.num_train_timesteps // self.num_inference_steps
        ts = [(self.num_inference_steps - 1 - i) * ratio for i in range(self.num_inference_steps)]
        self.timesteps = torch.tensor(ts, dtype=torch.int64, device=device)

    def scale_model_input(self, sample, timestep=None):
        return sample

    def coefficients(self, t: int) -> Tuple[float, float]:
        """(cx, ce) with prev_sample = cx * sample + ce * model_output  (eta = 0)."""
        prev_t = t - self.config.num_train_timesteps // self.num_inference_steps
        a_t = self._acp[t]
        a_p = self._acp[prev_t] if prev_t >= 0 else self.final_alpha_cumprod
        sa_t, sb_t = math.sqrt(a_t), math.sqrt(1.0 - a_t)
        sa_p, sb_p = math.sqrt(a_p), math.sqrt(1.0 - a_p)
        if self.config.prediction_type == "epsilon":
            # x0 = (x - sb_t e)/sa_t ; x' = sa_p x0 + sb_p e
            return sa_p / sa_t, sb_p - sa_p * sb_t / sa_t
        # v-prediction: x0 = sa_t x - sb_t v ; e = sa_t v + sb_t x
        return sa_p * sa_t + sb_p * sb_t, sb_p * sa_t - sa_p * sb_t

    def step(self, model_output, timestep, sample, eta: float = 0.0, **_):
        if eta != 0.0:
            raise NotImplementedError("leco_b200 DDIM implements eta=0 (the reference never passes eta)")
        cx, ce = self.coefficients(int(timestep))
        from . import ops
        # no host fallback: ops.axpby raises for CPU tensors / a missing CUDA library
        return SimpleNamespace(prev_sample=ops.axpby(sample, model_output, cx, ce))


def create_noise_scheduler(scheduler_name: str = "ddim", prediction_type: str = "epsilon"):
    """Same dispatch contract as model_util.create_noise_scheduler (model_util.py:230-278)."""
    name = scheduler_name.lower().replace(" ", "_")
    if name == "ddim":
        return DDIMScheduler(prediction_type)
    if name in ("ddpm", "lms", "euler_a"):
        raise NotImplementedError(f"scheduler '{name}' is outside the round-1 hot path (SURVEY §8f rank 2)")
    raise ValueError(f"Unknown scheduler name: {name}")
