"""`load_models` / `load_models_xl` with the reference's return tuples (model_util.py:104-129, :200-227) and the prompt
encoding of `train_util.py:60-130`, for an offline sandbox.

`pretrained_model_name_or_path` may be

  * a diffusers checkpoint DIRECTORY (`tokenizer/ text_encoder/ unet/` [+ `tokenizer_2/ text_encoder_2/` for SDXL]) —
    what the reference downloads from the hub (model_util.py:107-122).  Everything real is loaded: `ClipTokenizer`
    from `vocab.json`/`merges.txt`, `ClipTextEncoder` from `config.json` + weights, the UNet topology from
    `unet/config.json` and its weights from `diffusion_pytorch_model.safetensors|.bin` (the engine tree uses diffusers'
    parameter names, so keys match 1:1);
  * a single-file checkpoint in the ORIGINAL (LDM) key layout, `.ckpt` / `.safetensors` — what the reference passes to
    `from_single_file` (model_util.py:78-101, 173-197): UNet and text encoder(s) are converted key by key
    (`leco_b200.ckpt_convert`).  The file holds no vocabulary (diffusers fetches it from the hub): `tokenizer/`
    [+ `tokenizer_2/`] are read from the checkpoint's own directory or from `$LECO_TOKENIZER_DIR`;
  * a single .safetensors / .pt file holding a diffusers-format UNet state dict: UNet weights real, text side synthetic;
  * anything else (a hub name, an architecture name of `leco_b200.unet.SPECS`): seeded synthetic weights of the
    architecture the config flags select, and a stand-in text encoder that returns a seeded N(0,1) embedding per
    prompt string.  No checkpoint or vocabulary exists in this sandbox, so this is what every benchmark here runs;
    the run says so on stdout."""
from __future__ import annotations

import json
import os
from typing import List, Tuple

import torch

from .scheduler import create_noise_scheduler
from .synthetic import build_engine, prompt_embedding
from .text_encoder import ClipTextEncoder
from .tokenizer import ClipTokenizer
from .unet import SPECS, EngineUNet, UNetSpec


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


# ---- train_util.py:60-130 ---------------------------------------------------------------------------------------------
def text_tokenize(tokenizer: ClipTokenizer, prompts: List[str]) -> torch.Tensor:
    return tokenizer(prompts, padding="max_length", max_length=tokenizer.model_max_length, truncation=True,
                     return_tensors="pt").input_ids


def text_encode(text_encoder: ClipTextEncoder, tokens: torch.Tensor) -> torch.Tensor:
    return text_encoder(tokens.to(text_encoder.device))[0]


def encode_prompts(tokenizer, text_encoder, prompts: List[str]):
    """train_util.encode_prompts (train_util.py:77-86): [len(prompts), 77, D]."""
    if isinstance(text_encoder, SyntheticTextEncoder):
        embs = [text_encoder.encode(p) for p in prompts]
        return embs[0] if len(embs) == 1 else torch.cat(embs, 0)
    return text_encode(text_encoder, text_tokenize(tokenizer, prompts))


def text_encode_xl(text_encoder: ClipTextEncoder, tokens: torch.Tensor, num_images_per_prompt: int = 1):
    out = text_encoder(tokens.to(text_encoder.device), output_hidden_states=True)
    pooled = out[0]
    embeds = out.hidden_states[-2]                       # always the penultimate layer (train_util.py:98)
    b, s, _ = embeds.shape
    return embeds.repeat(1, num_images_per_prompt, 1).view(b * num_images_per_prompt, s, -1), pooled


def encode_prompts_xl(tokenizers, text_encoders, prompts: List[str], num_images_per_prompt: int = 1):
    """train_util.encode_prompts_xl (train_util.py:106-130): both encoders' penultimate states side by side and the
    SECOND encoder's projected pooled row -> `EmbedsXL`."""
    from .trainer import EmbedsXL
    if isinstance(text_encoders[0], SyntheticTextEncoder):
        assert len(prompts) == 1
        return text_encoders[0].encode(prompts[0])
    parts, pooled = [], None
    for tokenizer, text_encoder in zip(tokenizers, text_encoders):
        embeds, pooled = text_encode_xl(text_encoder, text_tokenize(tokenizer, prompts), num_images_per_prompt)
        parts.append(embeds)
    b = pooled.shape[0]
    pooled = pooled.repeat(1, num_images_per_prompt).view(b * num_images_per_prompt, -1)
    return EmbedsXL(torch.cat(parts, dim=-1), pooled)


# ---- UNet -------------------------------------------------------------------------------------------------------------
def unet_spec_from_config(cfg: dict, name: str = "checkpoint") -> UNetSpec:
    """`unet/config.json` (diffusers UNet2DConditionModel) -> the engine's topology record.  `attention_head_dim` is, by a
    long-standing diffusers naming slip, the NUMBER of heads per level unless `num_attention_heads` is given."""
    boc = tuple(cfg["block_out_channels"])
    levels = len(boc)

    def per_level(v, default):
        if v is None:
            v = default
        return tuple(v) if isinstance(v, (list, tuple)) else (v,) * levels
    heads = per_level(cfg.get("num_attention_heads") or cfg.get("attention_head_dim"), 8)
    text_time = cfg.get("addition_embed_type") == "text_time"
    kw = dict(block_out_channels=boc,
              attn_levels=tuple(t.startswith("CrossAttn") for t in cfg["down_block_types"]),
              layers_per_block=cfg.get("layers_per_block", 2), cross_attention_dim=cfg["cross_attention_dim"],
              num_heads=heads, transformer_depth=per_level(cfg.get("transformer_layers_per_block"), 1),
              use_linear_projection=bool(cfg.get("use_linear_projection", False)),
              norm_groups=cfg.get("norm_num_groups", 32), text_time=text_time,
              in_channels=cfg.get("in_channels", 4), out_channels=cfg.get("out_channels", 4))
    if text_time:
        kw.update(add_time_dim=cfg["addition_time_embed_dim"], add_proj_in=cfg["projection_class_embeddings_input_dim"])
    return UNetSpec(name, **kw)


def _read_weights(path_base: str):
    for ext in (".safetensors", ".bin", ".pt"):
        p = path_base + ext
        if os.path.isfile(p):
            if ext == ".safetensors":
                from safetensors.torch import load_file
                return load_file(p), p
            return torch.load(p, map_location="cpu"), p
    raise FileNotFoundError(f"no weights at {path_base}.safetensors|.bin|.pt")


def _engine_from_state(spec: UNetSpec, sd, device) -> EngineUNet:
    with torch.device(device):
        unet = EngineUNet(spec)
    unet.load_state_dict(dict(sd), strict=True)
    unet.requires_grad_(False)
    unet.eval()
    unet.pack(torch.device(device))
    return unet


def is_checkpoint_dir(path: str) -> bool:
    return bool(path) and os.path.isdir(path) and os.path.isfile(os.path.join(path, "unet", "config.json"))


def _load_unet(arch: str, name_or_path: str, device) -> Tuple[EngineUNet, str]:
    if is_checkpoint_dir(name_or_path):
        with open(os.path.join(name_or_path, "unet", "config.json"), encoding="utf-8") as f:
            spec = unet_spec_from_config(json.load(f), name=os.path.basename(os.path.normpath(name_or_path)))
        sd, where = _read_weights(os.path.join(name_or_path, "unet", "diffusion_pytorch_model"))
        return _engine_from_state(spec, sd, device), f"topology from unet/config.json, weights from {where}"
    if name_or_path and os.path.isfile(name_or_path):
        from . import ckpt_convert
        layout, sd, encs = ckpt_convert.split_single_file(ckpt_convert.read_single_file(name_or_path),
                                                          SPECS[arch].layers_per_block)
        _SINGLE_FILE_TEXT[os.path.abspath(name_or_path)] = encs       # consumed by _single_file_text_side (one read)
        return _engine_from_state(SPECS[arch], sd, device), f"weights from {name_or_path} ({layout} layout)"
    return build_engine(arch, device, seed=0), "synthetic seeded weights (no checkpoint available offline)"


_SINGLE_FILE_TEXT = {}


def _single_file_text_side(path: str, device, n_expected: int):
    """(tokenizers, encoders) of an LDM-layout single-file checkpoint, or None when the file carries no text encoder
    (a bare diffusers UNet).  model_util.py:78-101 / :173-197 get both from `from_single_file`; the vocabulary is not
    in the file, so `tokenizer/` [`tokenizer_2/`] must sit beside it or under $LECO_TOKENIZER_DIR."""
    from . import ckpt_convert
    encs = _SINGLE_FILE_TEXT.pop(os.path.abspath(path), None)
    if encs is None:
        _, _, encs = ckpt_convert.split_single_file(ckpt_convert.read_single_file(path))
    if not encs:
        return None
    if len(encs) != n_expected:
        raise ValueError(f"{path}: {len(encs)} text encoder(s) in the file, this loop needs {n_expected} "
                         f"({'load_models_xl' if n_expected == 2 else 'load_models'})")
    roots = [os.path.dirname(os.path.abspath(path))] + ([os.environ["LECO_TOKENIZER_DIR"]]
                                                        if os.environ.get("LECO_TOKENIZER_DIR") else [])
    toks, models = [], []
    for i, (sd, act, with_proj) in enumerate(encs):
        sub = "tokenizer" if i == 0 else f"tokenizer_{i + 1}"
        root = next((r for r in roots if os.path.isfile(os.path.join(r, sub, "vocab.json"))), None)
        if root is None:
            raise FileNotFoundError(f"single-file checkpoint {path}: no {sub}/vocab.json beside it or under "
                                    f"$LECO_TOKENIZER_DIR (the reference downloads the CLIP vocabulary from the hub)")
        toks.append(ClipTokenizer.from_pretrained(root, sub))
        enc = ClipTextEncoder(ckpt_convert.clip_spec_from_state(sd, f"{os.path.basename(path)}:{i}", act, with_proj))
        enc.load_state_dict(sd)
        models.append(enc.to(device))
    return toks, models


AVAILABLE_SCHEDULERS = ("ddim", "ddpm", "lms", "euler_a")     # model_util.py:23
DIFFUSERS_CACHE_DIR = None                                    # model_util.py:27 (nothing is downloaded here)
SDXL_TEXT_ENCODER_TYPE = ClipTextEncoder                      # model_util.py:25 (both towers are this class)


def load_diffusers_model(pretrained_model_name_or_path: str, v2: bool = False, clip_skip=None, weight_dtype=None,
                         device="cuda"):
    """model_util.py:30-75 -> (tokenizer, text_encoder, unet) from a diffusers checkpoint directory.  `clip_skip` is
    accepted like the reference's and, like there (`load_models` never passes it), unused by the training loops."""
    if not is_checkpoint_dir(pretrained_model_name_or_path):
        raise FileNotFoundError(f"{pretrained_model_name_or_path}: not a diffusers checkpoint directory (no hub access here)")
    tokenizer, text_encoder, unet, _ = load_models(pretrained_model_name_or_path, "ddim", v2=v2, device=device)
    return tokenizer, text_encoder, unet


def load_checkpoint_model(checkpoint_path: str, v2: bool = False, clip_skip=None, weight_dtype=None, device="cuda"):
    """model_util.py:78-101 -> (tokenizer, text_encoder, unet) from a single-file checkpoint (ckpt_convert.py)."""
    if not os.path.isfile(checkpoint_path):
        raise FileNotFoundError(checkpoint_path)
    tokenizer, text_encoder, unet, _ = load_models(checkpoint_path, "ddim", v2=v2, device=device)
    return tokenizer, text_encoder, unet


def load_diffusers_model_xl(pretrained_model_name_or_path: str, weight_dtype=None, device="cuda"):
    """model_util.py:132-170 -> (tokenizers, text_encoders, unet)."""
    if not is_checkpoint_dir(pretrained_model_name_or_path):
        raise FileNotFoundError(f"{pretrained_model_name_or_path}: not a diffusers checkpoint directory (no hub access here)")
    tokenizers, text_encoders, unet, _ = load_models_xl(pretrained_model_name_or_path, "ddim", device=device)
    return tokenizers, text_encoders, unet


def load_checkpoint_model_xl(checkpoint_path: str, weight_dtype=None, device="cuda"):
    """model_util.py:173-197 -> (tokenizers, text_encoders, unet) from a single-file SDXL checkpoint."""
    if not os.path.isfile(checkpoint_path):
        raise FileNotFoundError(checkpoint_path)
    tokenizers, text_encoders, unet, _ = load_models_xl(checkpoint_path, "ddim", device=device)
    return tokenizers, text_encoders, unet


def load_models(pretrained_model_name_or_path: str, scheduler_name: str = "ddim", v2: bool = False, v_pred: bool = False,
                weight_dtype=None, device="cuda", arch: str = None):
    """-> (tokenizer, text_encoder, unet, scheduler), model_util.py:104-129.  `weight_dtype` is accepted for signature
    parity: holders keep the checkpoint's dtype (the caller's `.to(device, dtype=...)` works as in train_lora.py:64-67),
    the kernel-layout copies are always bf16."""
    # an architecture name of leco_b200.unet.SPECS (e.g. the reduced-width "tiny21") selects synthetic weights of it
    arch = arch or (pretrained_model_name_or_path if pretrained_model_name_or_path in SPECS else ("sd21" if v2 else "sd15"))
    unet, how = _load_unet(arch, pretrained_model_name_or_path, device)
    if is_checkpoint_dir(pretrained_model_name_or_path):
        tokenizer = ClipTokenizer.from_pretrained(pretrained_model_name_or_path, "tokenizer")
        enc = ClipTextEncoder.from_pretrained(pretrained_model_name_or_path, "text_encoder", device=device)
        how += "; tokenizer + CLIP text encoder from the same directory"
    elif os.path.isfile(pretrained_model_name_or_path or "") and \
            (text := _single_file_text_side(pretrained_model_name_or_path, device, 1)) is not None:
        (tokenizer,), (enc,) = text
        how += "; CLIP text encoder converted from the same file, tokenizer from the directory beside it"
    else:
        tokenizer, enc = None, SyntheticTextEncoder(unet.spec.cross_attention_dim)
        how += "; stand-in text embeddings (no tokenizer / text encoder files)"
    print(f"leco_b200.load_models: {unet.spec.name} UNet, {how}")
    # model_util.py:124-127: v_pred only switches the scheduler's prediction type
    scheduler = create_noise_scheduler(scheduler_name, prediction_type="v_prediction" if v_pred else "epsilon")
    return tokenizer, enc, unet, scheduler


def load_models_xl(pretrained_model_name_or_path: str, scheduler_name: str = "ddim", weight_dtype=None, device="cuda",
                   arch: str = None):
    """-> (tokenizers, text_encoders, unet, scheduler), model_util.py:200-227."""
    arch = arch or (pretrained_model_name_or_path if pretrained_model_name_or_path in SPECS else "sdxl")
    unet, how = _load_unet(arch, pretrained_model_name_or_path, device)
    if is_checkpoint_dir(pretrained_model_name_or_path):
        d = pretrained_model_name_or_path
        tokenizers = [ClipTokenizer.from_pretrained(d, "tokenizer"), ClipTokenizer.from_pretrained(d, "tokenizer_2")]
        encoders = [ClipTextEncoder.from_pretrained(d, "text_encoder", device=device),
                    ClipTextEncoder.from_pretrained(d, "text_encoder_2", device=device)]
        how += "; both tokenizers + CLIP text encoders from the same directory"
    elif os.path.isfile(pretrained_model_name_or_path or "") and \
            (text := _single_file_text_side(pretrained_model_name_or_path, device, 2)) is not None:
        tokenizers, encoders = text
        how += "; both CLIP text encoders converted from the same file, tokenizers from the directory beside it"
    else:
        enc = SyntheticTextEncoder(unet.spec.cross_attention_dim, unet.spec.add_text_dim)
        tokenizers, encoders = [None, None], [enc, enc]
        how += "; stand-in text embeddings (no tokenizer / text encoder files)"
    print(f"leco_b200.load_models_xl: {unet.spec.name} UNet, {how}")
    return tokenizers, encoders, unet, create_noise_scheduler(scheduler_name)
