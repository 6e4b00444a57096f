"""`load_models` with the reference's return tuple (model_util.py:104-129, :200-227), for an offline sandbox.

The reference pulls tokenizer / CLIP text encoder / UNet from the Hugging Face hub or a checkpoint file (diffusers,
transformers).  Neither the libraries' weights nor a network exist here, so:
  * the UNet is a `leco_b200.unet.EngineUNet` of the architecture the config flags select (v2 -> SD2.1 layout, else
    SD1.5; XL through `load_models_xl`).  If `name_or_path` is a local .safetensors / .pt file holding a diffusers-format
    UNet state dict (the tree uses diffusers' parameter names, so keys match 1:1) it is loaded; otherwise seeded
    synthetic weights of that architecture are used and the run says so;
  * tokenizer / text encoder are a stand-in pair whose `encode_prompts` (train_util.py:60-130 surface) returns a seeded
    N(0,1) embedding per prompt string of the right shape.  SURVEY §8f rank 1 (real encoders) is not built."""
from __future__ import annotations

import os
from typing import Tuple

import torch

from .scheduler import create_noise_scheduler
from .synthetic import build_engine, prompt_embedding
from .unet import SPECS, EngineUNet


class SyntheticTextEncoder:
    """Stands for (tokenizer, text_encoder): deterministic [1,77,D] embedding per prompt (and a pooled [1,P] one for XL)."""

    def __init__(self, dim: int, pooled_dim: int = 0):
        self.dim, self.pooled_dim = dim, pooled_dim

    def to(self, *a, **k):
        return self

    def eval(self):
        return self

    def encode(self, prompt: str):
        text = prompt_embedding(prompt, self.dim)
        if not self.pooled_dim:
            return text
        from .trainer import EmbedsXL
        return EmbedsXL(text, prompt_embedding(prompt + "/pooled", self.pooled_dim)[0, :1])


def encode_prompts(tokenizer, text_encoder: SyntheticTextEncoder, prompts):
    """train_util.encode_prompts surface (train_util.py:96-104): one embedding per prompt, stacked on dim 0."""
    embs = [text_encoder.encode(p) for p in prompts]
    return embs[0] if len(embs) == 1 else torch.cat(embs, 0)


def _load_unet(arch: str, name_or_path: str, device) -> Tuple[EngineUNet, str]:
    if name_or_path and os.path.isfile(name_or_path):
        if name_or_path.endswith(".safetensors"):
            from safetensors.torch import load_file
            sd = load_file(name_or_path)
        else:
            sd = torch.load(name_or_path, map_location="cpu")
        with torch.device(device):
            unet = EngineUNet(SPECS[arch])
        unet.load_state_dict({k: v for k, v in sd.items()}, strict=True)
        unet.requires_grad_(False)
        unet.eval()
        unet.pack(torch.device(device))
        return unet, f"weights from {name_or_path}"
    return build_engine(arch, device, seed=0), "synthetic seeded weights (no checkpoint available offline)"


def load_models(pretrained_model_name_or_path: str, scheduler_name: str = "ddim", v2: bool = False, v_pred: bool = False,
                device="cuda", arch: str = None):
    """-> (tokenizer, text_encoder, unet, scheduler), model_util.py:104-129."""
    # an architecture name of leco_b200.unet.SPECS (e.g. the reduced-width "tiny21") selects synthetic weights of it
    arch = arch or (pretrained_model_name_or_path if pretrained_model_name_or_path in SPECS else ("sd21" if v2 else "sd15"))
    unet, how = _load_unet(arch, pretrained_model_name_or_path, device)
    print(f"leco_b200.load_models: {arch} UNet, {how}")
    enc = SyntheticTextEncoder(SPECS[arch].cross_attention_dim)
    # model_util.py:124-127: v_pred only switches the scheduler's prediction type
    scheduler = create_noise_scheduler(scheduler_name, prediction_type="v_prediction" if v_pred else "epsilon")
    return None, enc, unet, scheduler


def load_models_xl(pretrained_model_name_or_path: str, scheduler_name: str = "ddim", device="cuda", arch: str = None):
    """-> (tokenizers, text_encoders, unet, scheduler), model_util.py:200-227."""
    arch = arch or (pretrained_model_name_or_path if pretrained_model_name_or_path in SPECS else "sdxl")
    unet, how = _load_unet(arch, pretrained_model_name_or_path, device)
    print(f"leco_b200.load_models_xl: {arch} UNet, {how}")
    spec = SPECS[arch]
    enc = SyntheticTextEncoder(spec.cross_attention_dim, spec.add_text_dim)
    return [None, None], [enc, enc], unet, create_noise_scheduler(scheduler_name)
