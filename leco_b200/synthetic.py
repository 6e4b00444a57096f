"""Synthetic stand-ins for what the sandbox cannot provide (no checkpoints, no text encoder):
random-init weights of the right architecture and N(0,1) prompt embeddings.  Used by
bench.py / smoke(); parity tests load the oracle's seeded weights instead."""
from __future__ import annotations

import math
import zlib

import torch
import torch.nn as nn

from .unet import SPECS, EngineUNet


@torch.no_grad()
def init_weights_(model: nn.Module, seed: int = 0) -> nn.Module:
    """Variance-preserving random init (weights ~ N(0, 1/fan_in), residual-branch outputs damped,
    norm affine near identity) so activations stay O(1) through ~100 layers in bf16."""
    dev = next(model.parameters()).device
    g = torch.Generator(device=dev).manual_seed(seed)
    damped = ("conv2", "proj_out", "to_out.0", "ff.net.2")
    for name, m in model.named_modules():
        if isinstance(m, (nn.Linear, nn.Conv2d)):
            fan_in = m.weight[0].numel()
            std = (0.35 if name.endswith(damped) else 1.0) / math.sqrt(fan_in)
            m.weight.copy_(torch.randn(m.weight.shape, generator=g, device=dev) * std)
            if m.bias is not None:
                m.bias.copy_(torch.randn(m.bias.shape, generator=g, device=dev) * 0.05)
        elif isinstance(m, (nn.GroupNorm, nn.LayerNorm)):
            m.weight.copy_(1.0 + 0.1 * torch.randn(m.weight.shape, generator=g, device=dev))
            m.bias.copy_(0.1 * torch.randn(m.bias.shape, generator=g, device=dev))
    return model


def build_engine(arch: str, device="cuda", seed: int = 0) -> EngineUNet:
    with torch.device(device):
        eng = EngineUNet(SPECS[arch])
    init_weights_(eng, seed)
    eng.requires_grad_(False)
    eng.eval()
    eng.pack(torch.device(device))
    return eng


def prompt_embedding(prompt: str, dim: int) -> torch.Tensor:
    """[1,77,D] N(0,1) embedding seeded by the prompt text (stands in for CLIP, out of scope)."""
    g = torch.Generator(device="cpu").manual_seed(zlib.crc32(prompt.encode()) & 0x7FFFFFFF)
    return torch.randn((1, 77, dim), generator=g)
