"""Single-file Stable Diffusion checkpoints (`.ckpt` / `.safetensors` in the original LDM key layout) -> the
diffusers-layout state dicts the engine's parameter trees use.

The reference hands such a file to diffusers' `StableDiffusion[XL]Pipeline.from_single_file`
(`/root/reference/model_util.py:78-101, 173-197`, taken when `pretrained_model.name_or_path` ends in `.ckpt` /
`.safetensors`, `:111-118, 207-216`).  diffusers is a third-party dependency that is not vendored in the reference and
not installed here, so this file restates the PUBLISHED key layout of the two formats (parity unpinned, like
oracle/unet_ref.py: no golden file of the conversion exists offline; `tests/ldm_fixtures.py` writes the LDM layout
independently from the diffusers side and the round trip must be the identity):

UNet, prefix `model.diffusion_model.` (openaimodel.UNetModel -> UNet2DConditionModel)
    time_embed.0 / .2                      time_embedding.linear_1 / linear_2
    label_emb.0.0 / .0.2  (SDXL)           add_embedding.linear_1 / linear_2
    input_blocks.0.0                       conv_in
    input_blocks.i (i >= 1)                down block (i-1) // (L+1), layer (i-1) % (L+1)   [L = layers_per_block]
        .0 = ResBlock -> resnets.layer ; .1 = SpatialTransformer -> attentions.layer ; .0.op = downsamplers.0.conv
    middle_block.0 / .1 / .2               mid_block.resnets.0 / attentions.0 / resnets.1
    output_blocks.i                        up block i // (L+1), layer i % (L+1)
        .0 = ResBlock ; a later sub-module with `.conv` = upsamplers.0.conv, with `.norm` = attentions.layer
    out.0 / out.2                          conv_norm_out / conv_out
    ResBlock: in_layers.0 / .2 -> norm1 / conv1, emb_layers.1 -> time_emb_proj, out_layers.0 / .3 -> norm2 / conv2,
              skip_connection -> conv_shortcut.   SpatialTransformer keys are the same on both sides.
Text encoders
    SD1.x   `cond_stage_model.transformer.`          transformers CLIPTextModel keys as they are
    SD2.x   `cond_stage_model.model.`                open_clip layout (below); the LAST block is dropped (diffusers'
                                                     SD2 text_encoder has 23 layers and feeds the UNet its final-LN output)
    SDXL    `conditioner.embedders.0.transformer.`   CLIP ViT-L (transformers keys)
            `conditioner.embedders.1.model.`         open_clip bigG, all 32 blocks, `text_projection` transposed
    open_clip: positional_embedding, token_embedding.weight, ln_final, transformer.resblocks.N.{ln_1, ln_2,
    attn.in_proj_{weight,bias} (q|k|v stacked), attn.out_proj, mlp.c_fc, mlp.c_proj}, text_projection [D, P] (x @ P).
"""
from __future__ import annotations

import os
import re
from typing import Dict, Tuple

import torch

UNET_PREFIX = "model.diffusion_model."
_RES = {"in_layers.0": "norm1", "in_layers.2": "conv1", "emb_layers.1": "time_emb_proj", "out_layers.0": "norm2",
        "out_layers.3": "conv2", "skip_connection": "conv_shortcut"}


def read_single_file(path: str) -> Dict[str, torch.Tensor]:
    """`.safetensors` or a pickled `.ckpt` / `.pt` (optionally nested under "state_dict")."""
    if path.endswith(".safetensors"):
        from safetensors.torch import load_file
        return load_file(path)
    try:
        sd = torch.load(path, map_location="cpu", weights_only=True)
    except Exception as e:
        # old .ckpt files pickle training-framework objects next to the weights; unpickling those runs code
        if os.environ.get("LECO_TRUST_CKPT") != "1":
            raise RuntimeError(f"{path}: not a plain tensor pickle ({type(e).__name__}); set LECO_TRUST_CKPT=1 to unpickle "
                               f"it anyway (as the reference does), or convert it to .safetensors") from e
        sd = torch.load(path, map_location="cpu", weights_only=False)
    while isinstance(sd, dict) and "state_dict" in sd:
        sd = sd["state_dict"]
    return sd


def detect_layout(sd) -> str:
    """"ldm_sd1" / "ldm_sd2" / "ldm_sdxl" for original-layout checkpoints, "diffusers_unet" for a bare diffusers UNet."""
    keys = sd.keys()
    if any(k.startswith("conditioner.embedders.") for k in keys):
        return "ldm_sdxl"
    if any(k.startswith("cond_stage_model.model.") for k in keys):
        return "ldm_sd2"
    if any(k.startswith(UNET_PREFIX) for k in keys):
        return "ldm_sd1"
    if "conv_in.weight" in keys:
        return "diffusers_unet"
    raise ValueError("unrecognised checkpoint: neither an LDM-layout file (model.diffusion_model.*) nor a diffusers UNet")


def _res(key: str) -> str:
    for a, b in _RES.items():
        if key.startswith(a + "."):
            return b + key[len(a):]
    raise KeyError(f"unexpected ResBlock key {key!r}")


def convert_ldm_unet(sd, layers_per_block: int = 2) -> Dict[str, torch.Tensor]:
    """`model.diffusion_model.*` -> UNet2DConditionModel keys (tensors are shared, not copied)."""
    src = {k[len(UNET_PREFIX):]: v for k, v in sd.items() if k.startswith(UNET_PREFIX)}
    if not src:
        raise ValueError(f"no {UNET_PREFIX}* keys in the checkpoint")
    per = layers_per_block + 1
    # sub-module kind of every output_blocks.i.j (j >= 1): 'conv' -> upsampler, else attention
    up_kind = {}
    for k in src:
        m = re.match(r"output_blocks\.(\d+)\.(\d+)\.(\w+)", k)
        if m and int(m.group(2)) >= 1:
            ij = (int(m.group(1)), int(m.group(2)))
            if m.group(3) == "conv":
                up_kind[ij] = "up"
            else:
                up_kind.setdefault(ij, "attn")
    out = {}
    for k, v in src.items():
        head, _, rest = k.partition(".")
        if head == "time_embed":
            n, _, tail = rest.partition(".")
            out[f"time_embedding.linear_{1 + int(n) // 2}.{tail}"] = v
        elif head == "label_emb":
            m = re.match(r"0\.(\d+)\.(.+)", rest)
            out[f"add_embedding.linear_{1 + int(m.group(1)) // 2}.{m.group(2)}"] = v
        elif head == "out":
            n, _, tail = rest.partition(".")
            out[("conv_norm_out." if n == "0" else "conv_out.") + tail] = v
        elif head == "input_blocks":
            m = re.match(r"(\d+)\.(\d+)\.(.+)", rest)
            i, j, tail = int(m.group(1)), int(m.group(2)), m.group(3)
            if i == 0:
                out["conv_in." + tail] = v
                continue
            blk, layer = (i - 1) // per, (i - 1) % per
            if tail.startswith("op."):
                out[f"down_blocks.{blk}.downsamplers.0.conv.{tail[3:]}"] = v
            elif j == 0:
                out[f"down_blocks.{blk}.resnets.{layer}.{_res(tail)}"] = v
            else:
                out[f"down_blocks.{blk}.attentions.{layer}.{tail}"] = v
        elif head == "middle_block":
            n, _, tail = rest.partition(".")
            out[{"0": "mid_block.resnets.0.", "1": "mid_block.attentions.0.", "2": "mid_block.resnets.1."}[n]
                + (_res(tail) if n != "1" else tail)] = v
        elif head == "output_blocks":
            m = re.match(r"(\d+)\.(\d+)\.(.+)", rest)
            i, j, tail = int(m.group(1)), int(m.group(2)), m.group(3)
            blk, layer = i // per, i % per
            if j == 0:
                out[f"up_blocks.{blk}.resnets.{layer}.{_res(tail)}"] = v
            elif up_kind[(i, j)] == "up":
                out[f"up_blocks.{blk}.upsamplers.0.{tail}"] = v
            else:
                out[f"up_blocks.{blk}.attentions.{layer}.{tail}"] = v
        else:
            raise KeyError(f"unexpected UNet key {UNET_PREFIX}{k}")
    return out


def convert_hf_clip(sd, prefix: str) -> Dict[str, torch.Tensor]:
    """A transformers CLIPTextModel stored under `prefix` (SD1.x, SDXL embedder 0)."""
    out = {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}
    if out and not any(k.startswith("text_model.") for k in out):      # some exports drop the `text_model.` level
        out = {"text_model." + k: v for k, v in out.items()}
    return out


def convert_open_clip(sd, prefix: str, keep_layers: int = None) -> Dict[str, torch.Tensor]:
    """open_clip text tower under `prefix` -> transformers CLIPTextModel[WithProjection] keys.  `keep_layers`: number of
    leading transformer blocks to keep (SD2.x: all but the last)."""
    out = {}
    for k, v in sd.items():
        if not k.startswith(prefix):
            continue
        k = k[len(prefix):]
        if k == "positional_embedding":
            out["text_model.embeddings.position_embedding.weight"] = v
        elif k == "token_embedding.weight":
            out["text_model.embeddings.token_embedding.weight"] = v
        elif k.startswith("ln_final."):
            out["text_model.final_layer_norm." + k[len("ln_final."):]] = v
        elif k == "text_projection":
            out["text_projection.weight"] = v.t().contiguous()
        elif k.startswith("transformer.resblocks."):
            m = re.match(r"transformer\.resblocks\.(\d+)\.(.+)", k)
            n, tail = int(m.group(1)), m.group(2)
            if keep_layers is not None and n >= keep_layers:
                continue
            base = f"text_model.encoder.layers.{n}."
            if tail.startswith("attn.in_proj_"):
                kind = tail[len("attn.in_proj_"):]                    # weight | bias
                d = v.shape[0] // 3
                for j, name in enumerate(("q_proj", "k_proj", "v_proj")):
                    out[f"{base}self_attn.{name}.{kind}"] = v[j * d:(j + 1) * d].contiguous()
            else:
                for a, b in (("ln_1.", "layer_norm1."), ("ln_2.", "layer_norm2."), ("attn.out_proj.", "self_attn.out_proj."),
                             ("mlp.c_fc.", "mlp.fc1."), ("mlp.c_proj.", "mlp.fc2.")):
                    if tail.startswith(a):
                        out[base + b + tail[len(a):]] = v
                        break
                else:
                    raise KeyError(f"unexpected open_clip key {prefix}{k}")
        elif k in ("logit_scale", "attn_mask") or k.startswith("visual."):
            continue
        else:
            raise KeyError(f"unexpected open_clip key {prefix}{k}")
    return out


def clip_spec_from_state(sd, name: str, hidden_act: str, with_projection: bool = False):
    """The CLIPTextConfig values the weights themselves determine (head width 64 in every SD text tower)."""
    from .text_encoder import ClipTextSpec
    tok = sd["text_model.embeddings.token_embedding.weight"]
    layers = 1 + max(int(m.group(1)) for m in (re.match(r"text_model\.encoder\.layers\.(\d+)\.", k) for k in sd) if m)
    hidden = tok.shape[1]
    return ClipTextSpec(name, vocab_size=tok.shape[0], hidden_size=hidden,
                        intermediate_size=sd["text_model.encoder.layers.0.mlp.fc1.weight"].shape[0],
                        num_hidden_layers=layers, num_attention_heads=max(1, hidden // 64),
                        max_position_embeddings=sd["text_model.embeddings.position_embedding.weight"].shape[0],
                        hidden_act=hidden_act,
                        projection_dim=sd["text_projection.weight"].shape[0] if with_projection else 0,
                        eos_token_id=2)    # the published configs carry the legacy id: pooled row = argmax of the ids


def split_single_file(sd, layers_per_block: int = 2) -> Tuple[str, Dict[str, torch.Tensor], list]:
    """-> (layout, UNet state dict, [(text-encoder state dict, hidden_act, with_projection), ...])."""
    layout = detect_layout(sd)
    if layout == "diffusers_unet":
        return layout, dict(sd), []
    unet = convert_ldm_unet(sd, layers_per_block)
    if layout == "ldm_sd1":
        encs = [(convert_hf_clip(sd, "cond_stage_model.transformer."), "quick_gelu", False)]
    elif layout == "ldm_sd2":
        n = 1 + max(int(m.group(1)) for m in
                    (re.match(r"cond_stage_model\.model\.transformer\.resblocks\.(\d+)\.", k) for k in sd) if m)
        te = convert_open_clip(sd, "cond_stage_model.model.", keep_layers=n - 1)
        te.pop("text_projection.weight", None)                         # CLIPTextModel: no projection head
        encs = [(te, "gelu", False)]
    else:
        encs = [(convert_hf_clip(sd, "conditioner.embedders.0.transformer."), "quick_gelu", False),
                (convert_open_clip(sd, "conditioner.embedders.1.model."), "gelu", True)]
    return layout, unet, [e for e in encs if e[0]]
