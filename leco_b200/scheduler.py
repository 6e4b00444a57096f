"""Noise schedulers with the diffusers call surface the reference uses.

Mirrors what `model_util.create_noise_scheduler(name, prediction_type)` returns (model_util.py:230-278: DDIM / DDPM with
scaled-linear betas 0.00085..0.012, 1000 train steps, clip_sample=False; LMSDiscrete / EulerAncestralDiscrete with the
same betas) for the calls on the hot path:
  set_timesteps / timesteps        train_lora.py:143-145, 195-199
  scale_model_input                train_util.py:153      (identity for DDIM/DDPM, 1/sqrt(sigma^2+1) for LMS / Euler-a)
  step(...).prev_sample            train_util.py:190
  init_noise_sigma                 train_util.py:55       (1.0 for DDIM/DDPM, max sigma for LMS / Euler-a)
Every update is  x' = cx*x + ce*model_output + cn*noise + sum_j l_j * d_{-j}  with d = dx*x + dg*model_output (the
LMS derivative history); `plan(i)` returns that coefficient row for step i of the current grid.  The fused trainer puts
the rows of the whole grid in a device table and folds the update into the CFG-combine kernel (ops.guided_step for DDIM,
ops.sched_step for the others); `step()` applies the same row for drop-in callers (train_util.diffusion).
Coefficient row layout (fp32[12]): [guidance, cx, ce, cn, in_scale, dx, dg, l0, l1, l2, l3, history_slot].
"""
from __future__ import annotations

import math
from types import SimpleNamespace
from typing import List, Tuple

import torch

ROW = 12


def _alphas_cumprod(n=1000, beta_start=0.00085, beta_end=0.012) -> List[float]:
    # float32 table exactly like torch.linspace(...)**2 -> cumprod (diffusers builds it in fp32)
    betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, n, dtype=torch.float32) ** 2
    return torch.cumprod(1.0 - betas, dim=0).double().tolist()


class _Scheduler:
    init_noise_sigma = 1.0
    order = 1
    needs_noise = False       # the step draws Gaussian noise (DDPM, Euler-a)
    history = 0               # derivative history depth (LMS: 4)

    def __init__(self, prediction_type: str = "epsilon", num_train_timesteps: int = 1000):
        if prediction_type not in ("epsilon", "v_prediction"):
            raise ValueError(f"unsupported prediction_type {prediction_type}")
        self.config = SimpleNamespace(prediction_type=prediction_type, num_train_timesteps=num_train_timesteps)
        self._acp = _alphas_cumprod(num_train_timesteps)
        self.alphas_cumprod = torch.tensor(self._acp, dtype=torch.float32)
        self.final_alpha_cumprod = 1.0
        self.num_inference_steps = None
        self.timesteps = torch.arange(num_train_timesteps - 1, -1, -1, dtype=torch.int64)
        self._ts = list(range(num_train_timesteps - 1, -1, -1))
        self._hist = []

    # ---- grid
    def set_timesteps(self, num_inference_steps: int, device=None):
        """'leading' spacing, steps_offset 0:  t_i = (n-1-i) * (T // n)   (DDIM / DDPM defaults)."""
        self.num_inference_steps = int(num_inference_steps)
        ratio = self.config.num_train_timesteps // self.num_inference_steps
        ts = [(self.num_inference_steps - 1 - i) * ratio for i in range(self.num_inference_steps)]
        self.timesteps = torch.tensor(ts, dtype=torch.int64, device=device)
        self._ts = ts

    def _index(self, timestep) -> int:
        t = float(timestep)
        for i, v in enumerate(self._ts):
            if float(v) == t:
                return i
        raise ValueError(f"timestep {t} is not on the current grid")

    def in_scale(self, timestep) -> float:
        return 1.0

    def in_scale_at_train_timestep(self, t: int) -> float:
        """UNet input scale at integer training timestep t (the 1000-step grid of train_lora.py:195-199)."""
        return 1.0

    def scale_model_input(self, sample, timestep=None):
        s = self.in_scale(timestep)
        if s == 1.0:
            return sample
        from . import ops
        return ops.axpby(sample, sample, s, 0.0)

    # ---- one step
    def plan(self, i: int) -> List[float]:
        raise NotImplementedError

    def step(self, model_output, timestep, sample, eta: float = 0.0, generator=None, noise=None, **_):
        """prev_sample on the device through the library's kernels (no host fallback: CPU tensors raise)."""
        if eta != 0.0:
            raise NotImplementedError("leco_b200 schedulers implement eta=0 (the reference never passes eta)")
        from . import ops
        i = self._index(timestep)
        row = self.plan(i)
        _, cx, ce, cn, _, dx, dg, l0, l1, l2, l3, _ = row
        prev = ops.axpby(sample, model_output, cx, ce)
        if self.needs_noise and cn != 0.0:
            if noise is None:   # diffusers: randn_tensor(model_output.shape, device=model_output.device, generator=generator)
                noise = torch.randn(model_output.shape, generator=generator, device=model_output.device,
                                    dtype=model_output.dtype)
            prev = ops.axpby(prev, noise.to(prev.dtype), 1.0, cn)
        if self.history:
            d = ops.axpby(sample, model_output, dx, dg)
            self._hist.append(d)
            if len(self._hist) > self.history:
                self._hist.pop(0)
            for c, h in zip((l0, l1, l2, l3), reversed(self._hist)):
                if c != 0.0:
                    prev = ops.axpby(prev, h, 1.0, c)
        return SimpleNamespace(prev_sample=prev)

    def table(self, guidance: float) -> List[List[float]]:
        """coefficient rows of every step of the current grid (guidance in slot 0)."""
        rows = []
        for i in range(self.num_inference_steps):
            r = self.plan(i)
            r[0] = guidance
            rows.append(r)
        return rows


class DDIMScheduler(_Scheduler):
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

    def plan(self, i: int) -> List[float]:
        cx, ce = self.coefficients(int(self._ts[i]))
        return [0.0, cx, ce, 0.0, 1.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0]


class DDPMScheduler(_Scheduler):
    """scheduling_ddpm.py (diffusers 0.20.0) as built at model_util.py:247-256: variance_type "fixed_small", no clipping;
    prev = x0_coeff*x0 + x_coeff*x + sqrt(var)*noise for t > 0."""
    needs_noise = True

    def plan(self, i: int) -> List[float]:
        t = int(self._ts[i])
        n_inf = self.num_inference_steps or self.config.num_train_timesteps
        prev_t = t - self.config.num_train_timesteps // n_inf
        a_t = self._acp[t]
        a_p = self._acp[prev_t] if prev_t >= 0 else 1.0
        b_t, b_p = 1.0 - a_t, 1.0 - a_p
        cur_alpha = a_t / a_p
        cur_beta = 1.0 - cur_alpha
        x0_c = math.sqrt(a_p) * cur_beta / b_t
        x_c = math.sqrt(cur_alpha) * b_p / b_t
        if self.config.prediction_type == "epsilon":      # x0 = (x - sqrt(b_t) e) / sqrt(a_t)
            cx, ce = x0_c / math.sqrt(a_t) + x_c, -x0_c * math.sqrt(b_t) / math.sqrt(a_t)
        else:                                             # x0 = sqrt(a_t) x - sqrt(b_t) v
            cx, ce = x0_c * math.sqrt(a_t) + x_c, -x0_c * math.sqrt(b_t)
        cn = math.sqrt(max(b_p / b_t * cur_beta, 1e-20)) if t > 0 else 0.0
        return [0.0, cx, ce, cn, 1.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0]


class _SigmaScheduler(_Scheduler):
    """k-diffusion style schedulers (scheduling_lms_discrete.py / scheduling_euler_ancestral_discrete.py, diffusers
    0.20.0): sigma = sqrt((1-abar)/abar), "linspace" timestep spacing with linearly interpolated sigmas, UNet input
    scaled by 1/sqrt(sigma^2+1), init_noise_sigma = max sigma."""

    def __init__(self, prediction_type: str = "epsilon", num_train_timesteps: int = 1000):
        super().__init__(prediction_type, num_train_timesteps)
        acp32 = self.alphas_cumprod
        self._train_sigmas = (((1 - acp32) / acp32) ** 0.5).double().tolist()      # fp32 values, like diffusers' table
        self.init_noise_sigma = max(self._train_sigmas)
        self._ts = [float(t) for t in range(num_train_timesteps - 1, -1, -1)]
        self._sig = self._train_sigmas[::-1] + [0.0]
        self.sigmas = torch.tensor(self._sig, dtype=torch.float32)
        self.timesteps = torch.tensor(self._ts, dtype=torch.float64)

    def set_timesteps(self, num_inference_steps: int, device=None):
        n = self.num_inference_steps = int(num_inference_steps)
        T = self.config.num_train_timesteps
        import numpy as np
        ts = np.linspace(0, T - 1, n, dtype=float)[::-1].copy()
        sig = np.interp(ts, np.arange(0, T), np.array(self._train_sigmas))
        sig32 = np.concatenate([sig, [0.0]]).astype(np.float32)                        # diffusers keeps the table in fp32
        self._ts = [float(t) for t in ts]
        self._sig = [float(s) for s in sig32]
        self.sigmas = torch.from_numpy(sig32).to(device)
        self.timesteps = torch.from_numpy(ts).to(device)
        self._hist = []

    def in_scale(self, timestep) -> float:
        s = self._sig[self._index(timestep)]
        return 1.0 / math.sqrt(s * s + 1.0)

    def in_scale_at_train_timestep(self, t: int) -> float:
        s = float(torch.tensor(self._train_sigmas[int(t)], dtype=torch.float32))
        return 1.0 / math.sqrt(s * s + 1.0)

    def _derivative(self, sigma: float) -> Tuple[float, float]:
        """d = (x - x0)/sigma = dx*x + dg*model_output."""
        if self.config.prediction_type == "epsilon":     # x0 = x - sigma e
            return 0.0, 1.0
        # v-prediction: x0 = -sigma/sqrt(sigma^2+1) v + x/(sigma^2+1)
        return sigma / (sigma * sigma + 1.0), 1.0 / math.sqrt(sigma * sigma + 1.0)


class EulerAncestralDiscreteScheduler(_SigmaScheduler):
    needs_noise = True

    def plan(self, i: int) -> List[float]:
        s_from, s_to = self._sig[i], self._sig[i + 1]
        s_up = math.sqrt(max(s_to ** 2 * (s_from ** 2 - s_to ** 2) / s_from ** 2, 0.0))
        s_down = math.sqrt(max(s_to ** 2 - s_up ** 2, 0.0))
        dx, dg = self._derivative(s_from)
        dt = s_down - s_from
        return [0.0, 1.0 + dt * dx, dt * dg, s_up, 1.0 / math.sqrt(s_from ** 2 + 1.0), 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0]


class LMSDiscreteScheduler(_SigmaScheduler):
    history = 4

    def lms_coefficient(self, order: int, t: int, current_order: int) -> float:
        """integral over [sigma_t, sigma_{t+1}] of the Lagrange basis polynomial of `current_order` (the reference
        integrates it numerically, scipy.integrate.quad epsrel 1e-4; for a polynomial of degree <= 3 Gauss-Kronrod is
        exact to rounding, so the closed-form integral is the same number)."""
        import numpy as np
        poly = np.poly1d([1.0])
        for k in range(order):
            if k == current_order:
                continue
            denom = self._sig[t - current_order] - self._sig[t - k]
            poly = poly * np.poly1d([1.0 / denom, -self._sig[t - k] / denom])
        P = poly.integ()
        return float(P(self._sig[t + 1]) - P(self._sig[t]))

    def plan(self, i: int) -> List[float]:
        s = self._sig[i]
        dx, dg = self._derivative(s)
        order = min(i + 1, 4)
        l = [self.lms_coefficient(order, i, o) for o in range(order)] + [0.0] * (4 - order)
        return [0.0, 1.0, 0.0, 0.0, 1.0 / math.sqrt(s * s + 1.0), dx, dg, l[0], l[1], l[2], l[3], float(i % 4)]


def create_noise_scheduler(scheduler_name: str = "ddpm", prediction_type: str = "epsilon"):
    """Same dispatch contract as model_util.create_noise_scheduler (model_util.py:230-278)."""
    name = scheduler_name.lower().replace(" ", "_")
    if name == "ddim":
        return DDIMScheduler(prediction_type)
    if name == "ddpm":
        return DDPMScheduler(prediction_type)
    if name == "lms":
        return LMSDiscreteScheduler(prediction_type)
    if name == "euler_a":
        return EulerAncestralDiscreteScheduler(prediction_type)
    raise ValueError(f"Unknown scheduler name: {name}")
