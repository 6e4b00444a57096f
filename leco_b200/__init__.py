"""leco_b200 — B200-native (sm_100a) implementation of the LECO training-step hot path.

Product path = hand-written CUDA kernels behind the C ABI in include/leco_b200.h; this package
is the host-side mirror of the reference's call surface (LoRANetwork / predict_noise / the
train_lora.py loop body).  There is no CPU or PyTorch fallback: without the built library
(`python -m leco_b200.build`) and a CUDA device the compute entry points raise.
"""
from .unet import SPECS, EngineUNet, UNetSpec  # noqa: F401
from .lora import LoRANetwork, LoRAModule, FlatAdamW, FlatOptimizer  # noqa: F401
from .scheduler import (DDIMScheduler, DDPMScheduler, EulerAncestralDiscreteScheduler, LMSDiscreteScheduler,  # noqa: F401
                        create_noise_scheduler)

__all__ = ["SPECS", "EngineUNet", "UNetSpec", "LoRANetwork", "LoRAModule", "FlatAdamW", "FlatOptimizer", "DDIMScheduler",
           "DDPMScheduler", "LMSDiscreteScheduler", "EulerAncestralDiscreteScheduler", "create_noise_scheduler"]
