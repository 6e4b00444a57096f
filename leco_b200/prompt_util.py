"""The reference's prompt_util.py names on the engine's types, for code written against the reference
(`from prompt_util import PromptEmbedsPair, PromptEmbedsXL, PromptEmbedsCache, PromptSettings, load_prompts_from_yaml`).

`PromptEmbedsPair` IS a `leco_b200.trainer.PromptPair` (so `LecoTrainer` takes it as is) built with the reference's
constructor signature (prompt_util.py:86-107) and carrying the reference's host-side objective `loss(**kwargs)`
(:107-148) for the drop-in loop, where the four predictions are CPU tensors as in train_lora.py:265-270.  The fused
trainer computes the same objective on the device (`leco_loss` kernel) and never calls it."""
from __future__ import annotations

from typing import Optional, Union

import torch

from .config_util import ACTION_TYPES, PromptSettings, load_prompts_from_yaml  # noqa: F401  (same names as the reference)
from .trainer import EmbedsXL, PromptPair


class PromptEmbedsXL(EmbedsXL):
    """prompt_util.py:17-23: positional (text_embeds, pooled_embeds)."""

    def __init__(self, *args) -> None:
        super().__init__(args[0], args[1])


PROMPT_EMBEDDING = Union[torch.Tensor, PromptEmbedsXL]


class PromptEmbedsCache:
    """prompt_util.py:30-40.  NB the reference declares `prompts` on the CLASS, so every instance shares one dict; kept."""
    prompts: dict = {}

    def __setitem__(self, name: str, value) -> None:
        self.prompts[name] = value

    def __getitem__(self, name: str) -> Optional[object]:
        return self.prompts.get(name)


class PromptEmbedsPair(PromptPair):
    def __init__(self, loss_fn, target, positive, unconditional, neutral, settings) -> None:
        super().__init__(target=target, positive=positive, unconditional=unconditional, neutral=neutral,
                         guidance_scale=settings.guidance_scale, resolution=settings.resolution,
                         dynamic_resolution=settings.dynamic_resolution, batch_size=settings.batch_size,
                         action=settings.action, dynamic_crops=getattr(settings, "dynamic_crops", False))
        self.loss_fn = loss_fn

    def _erase(self, target_latents, positive_latents, unconditional_latents, neutral_latents):
        """prompt_util.py:109-122: the target prediction should lose the positive concept."""
        return self.loss_fn(target_latents,
                            neutral_latents - self.guidance_scale * (positive_latents - unconditional_latents))

    def _enhance(self, target_latents, positive_latents, unconditional_latents, neutral_latents):
        """prompt_util.py:124-137: the target prediction should gain the positive concept."""
        return self.loss_fn(target_latents,
                            neutral_latents + self.guidance_scale * (positive_latents - unconditional_latents))

    def loss(self, **kwargs):
        if self.action == "erase":
            return self._erase(**kwargs)
        if self.action == "enhance":
            return self._enhance(**kwargs)
        raise ValueError("action must be erase or enhance")
