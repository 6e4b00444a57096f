"""CLIP text encoder on the engine's kernels: the prompt-encoding prologue.

The reference encodes every distinct prompt once before the loop (`/root/reference/train_lora.py:106-132`) through
transformers' `CLIPTextModel` / `CLIPTextModelWithProjection`:

    train_util.text_encode      (train_util.py:73-74)   text_encoder(tokens)[0]                    -> last_hidden_state
    train_util.text_encode_xl   (train_util.py:88-103)  text_encoder(tokens, output_hidden_states=True)
                                                        [0] = pooled/projected, .hidden_states[-2]  -> penultimate layer

This module is that call surface.  The parameter tree carries transformers' names (`text_model.embeddings.*`,
`text_model.encoder.layers.N.{layer_norm1,self_attn.{q,k,v,out}_proj,layer_norm2,mlp.{fc1,fc2}}`,
`text_model.final_layer_norm`, `text_projection`) so a checkpoint's `text_encoder/` state dict loads key for key.
The arithmetic (transformers `modeling_clip.py`, CLIPTextTransformer): token + position embedding; per layer
pre-LayerNorm causal self-attention (scale d^-1/2) and a quick_gelu / erf-GELU MLP, both residual; final LayerNorm;
pooled row = the first <|endoftext|> position (legacy eos_token_id 2: argmax of the ids) of the final-LayerNorm
output, times `text_projection` when the model has one.

Kernels: `leco_embed_tokens`, `leco_layer_norm`, `leco_gemm_bf16` (q/k/v as ONE [3D, D] GEMM with bias; bias + residual
epilogues), `leco_gemm_batched` + `leco_softmax_rows_causal` (77 x 77 scores: one tile, nothing to stream, so the
materialised attention is used rather than the flash kernel), `leco_activation`.  bf16 activations, fp32 accumulation.
The backend is injectable like `EngineUNet`'s: the product runs `leco_b200.ops` (CUDA, no fallback); the CPU suite
injects a plain-torch double to check this file's wiring against transformers' own model.
"""
from __future__ import annotations

import json
import os
from dataclasses import dataclass
from typing import Dict, List, Optional

import torch
import torch.nn as nn

ACT_KINDS = {"quick_gelu": 1, "gelu": 2}


@dataclass
class ClipTextSpec:
    """The `text_encoder/config.json` values (transformers CLIPTextConfig) that define the network."""
    name: str
    vocab_size: int = 49408
    hidden_size: int = 768
    intermediate_size: int = 3072
    num_hidden_layers: int = 12
    num_attention_heads: int = 12
    max_position_embeddings: int = 77
    hidden_act: str = "quick_gelu"
    layer_norm_eps: float = 1e-5
    projection_dim: int = 0            # > 0: CLIPTextModelWithProjection (SDXL text_encoder_2)
    eos_token_id: int = 49407          # 2 = the legacy config value: pooled row = argmax(ids)

    @classmethod
    def from_config(cls, cfg: dict, name: str = "config") -> "ClipTextSpec":
        with_proj = "CLIPTextModelWithProjection" in (cfg.get("architectures") or [])
        return cls(name, vocab_size=cfg["vocab_size"], hidden_size=cfg["hidden_size"],
                   intermediate_size=cfg["intermediate_size"], num_hidden_layers=cfg["num_hidden_layers"],
                   num_attention_heads=cfg["num_attention_heads"],
                   max_position_embeddings=cfg.get("max_position_embeddings", 77),
                   hidden_act=cfg.get("hidden_act", "quick_gelu"), layer_norm_eps=cfg.get("layer_norm_eps", 1e-5),
                   projection_dim=cfg.get("projection_dim", 0) if with_proj else 0,
                   eos_token_id=cfg.get("eos_token_id", 2))


TEXT_SPECS: Dict[str, ClipTextSpec] = {
    # SD1.x text_encoder and SDXL text_encoder (OpenAI CLIP ViT-L/14)
    "clip_l": ClipTextSpec("clip_l"),
    # SD2.x text_encoder (OpenCLIP ViT-H/14 with the last layer already dropped by the checkpoint: 23 layers)
    "openclip_h": ClipTextSpec("openclip_h", hidden_size=1024, intermediate_size=4096, num_hidden_layers=23,
                               num_attention_heads=16, hidden_act="gelu"),
    # SDXL text_encoder_2 (OpenCLIP ViT-bigG/14, with text projection)
    "openclip_bigg": ClipTextSpec("openclip_bigg", hidden_size=1280, intermediate_size=5120, num_hidden_layers=32,
                                  num_attention_heads=20, hidden_act="gelu", projection_dim=1280),
    # reduced twins for the parity suite
    "tiny_clip": ClipTextSpec("tiny_clip", vocab_size=600, hidden_size=64, intermediate_size=128, num_hidden_layers=3,
                              num_attention_heads=2, eos_token_id=599),
    "tiny_clip_proj": ClipTextSpec("tiny_clip_proj", vocab_size=600, hidden_size=128, intermediate_size=256,
                                   num_hidden_layers=2, num_attention_heads=4, hidden_act="gelu", projection_dim=64,
                                   eos_token_id=2),
}


class _Holder(nn.Module):
    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("parameter holder; call ClipTextEncoder")


class _Embeddings(_Holder):
    def __init__(self, s: ClipTextSpec):
        super().__init__()
        self.token_embedding = nn.Embedding(s.vocab_size, s.hidden_size)
        self.position_embedding = nn.Embedding(s.max_position_embeddings, s.hidden_size)


class _SelfAttn(_Holder):
    def __init__(self, d: int):
        super().__init__()
        self.k_proj, self.v_proj, self.q_proj, self.out_proj = (nn.Linear(d, d) for _ in range(4))


class _Mlp(_Holder):
    def __init__(self, d: int, inner: int):
        super().__init__()
        self.fc1, self.fc2 = nn.Linear(d, inner), nn.Linear(inner, d)


class _Layer(_Holder):
    def __init__(self, s: ClipTextSpec):
        super().__init__()
        self.self_attn = _SelfAttn(s.hidden_size)
        self.layer_norm1 = nn.LayerNorm(s.hidden_size, eps=s.layer_norm_eps)
        self.mlp = _Mlp(s.hidden_size, s.intermediate_size)
        self.layer_norm2 = nn.LayerNorm(s.hidden_size, eps=s.layer_norm_eps)


class _Encoder(_Holder):
    def __init__(self, s: ClipTextSpec):
        super().__init__()
        self.layers = nn.ModuleList([_Layer(s) for _ in range(s.num_hidden_layers)])


class _TextModel(_Holder):
    def __init__(self, s: ClipTextSpec):
        super().__init__()
        self.embeddings = _Embeddings(s)
        self.encoder = _Encoder(s)
        self.final_layer_norm = nn.LayerNorm(s.hidden_size, eps=s.layer_norm_eps)


class ClipOutput(tuple):
    """`out[0]`, `out.hidden_states`, `out.last_hidden_state`, `out.text_embeds` as the reference reads them."""
    last_hidden_state: torch.Tensor
    pooler_output: torch.Tensor
    text_embeds: Optional[torch.Tensor]
    hidden_states: Optional[tuple]


def _output(first, last, pooled, text_embeds, hidden_states):
    o = ClipOutput((first,))
    o.last_hidden_state, o.pooler_output, o.text_embeds, o.hidden_states = last, pooled, text_embeds, hidden_states
    return o


class ClipTextEncoder(nn.Module):
    def __init__(self, spec: ClipTextSpec, backend=None):
        super().__init__()
        self.spec = spec
        self.text_model = _TextModel(spec)
        if spec.projection_dim:
            self.text_projection = nn.Linear(spec.hidden_size, spec.projection_dim, bias=False)
        self._backend = backend
        self._pack = None
        self.compute_dtype = torch.bfloat16     # the kernels' type; the CPU wiring tests run the double in fp32
        self.requires_grad_(False)

    # ---- weights ---------------------------------------------------------------------------------------------------
    @classmethod
    def from_pretrained(cls, directory: str, subfolder: str = "text_encoder", device="cuda") -> "ClipTextEncoder":
        """A diffusers checkpoint's `text_encoder[_2]/` folder: config.json + model.safetensors | pytorch_model.bin."""
        d = os.path.join(directory, subfolder) if subfolder else directory
        with open(os.path.join(d, "config.json"), encoding="utf-8") as f:
            spec = ClipTextSpec.from_config(json.load(f), name=os.path.basename(os.path.normpath(d)))
        st_path, pt_path = os.path.join(d, "model.safetensors"), os.path.join(d, "pytorch_model.bin")
        if os.path.isfile(st_path):
            from safetensors.torch import load_file
            sd = load_file(st_path)
        elif os.path.isfile(pt_path):
            sd = torch.load(pt_path, map_location="cpu")
        else:
            raise FileNotFoundError(f"no model.safetensors / pytorch_model.bin under {d}")
        enc = cls(spec)
        enc.load_state_dict(sd)
        return enc.to(device)

    def load_state_dict(self, state_dict, strict: bool = True, **k):
        # older transformers saved the position_ids buffer; it is not a weight
        sd = {key: v for key, v in state_dict.items() if not key.endswith("position_ids")}
        self._pack = None
        return super().load_state_dict(sd, strict=strict, **k)

    def _apply(self, fn, *a, **k):
        self._pack = None
        return super()._apply(fn, *a, **k)

    @property
    def device(self):
        return self.text_model.final_layer_norm.weight.device

    def backend(self):
        if self._backend is None:
            from . import ops
            self._backend = ops
        return self._backend

    def _packed(self, dev):
        """bf16 kernel-layout copies: q/k/v stacked to one [3D, D] weight and [3D] bias."""
        if self._pack is not None and self._pack["device"] == dev and self._pack["dtype"] == self.compute_dtype:
            return self._pack
        bf = self.compute_dtype

        def c(t):
            return t.detach().to(dev, bf).contiguous()
        tm = self.text_model
        layers = []
        for l in tm.encoder.layers:
            a = l.self_attn
            layers.append(dict(
                ln1=(c(l.layer_norm1.weight), c(l.layer_norm1.bias)),
                wqkv=c(torch.cat([a.q_proj.weight, a.k_proj.weight, a.v_proj.weight], 0)),
                bqkv=c(torch.cat([a.q_proj.bias, a.k_proj.bias, a.v_proj.bias], 0)),
                wo=c(a.out_proj.weight), bo=c(a.out_proj.bias),
                ln2=(c(l.layer_norm2.weight), c(l.layer_norm2.bias)),
                w1=c(l.mlp.fc1.weight), b1=c(l.mlp.fc1.bias), w2=c(l.mlp.fc2.weight), b2=c(l.mlp.fc2.bias)))
        self._pack = dict(device=dev, dtype=bf, tok=c(tm.embeddings.token_embedding.weight),
                          pos=c(tm.embeddings.position_embedding.weight), layers=layers,
                          lnf=(c(tm.final_layer_norm.weight), c(tm.final_layer_norm.bias)),
                          proj=c(self.text_projection.weight) if self.spec.projection_dim else None)
        return self._pack

    # ---- forward ---------------------------------------------------------------------------------------------------
    def eos_rows(self, input_ids: torch.Tensor) -> torch.Tensor:
        """Row of the pooled token per prompt (CLIPTextTransformer.forward: legacy argmax, else first eos id)."""
        ids = input_ids.to("cpu", torch.long)
        if self.spec.eos_token_id == 2:
            return ids.argmax(dim=-1)
        return (ids == self.spec.eos_token_id).int().argmax(dim=-1)

    @torch.no_grad()
    def forward(self, input_ids: torch.Tensor, output_hidden_states: bool = False, **_):
        s, be = self.spec, self.backend()
        dev = self.device
        if input_ids.dim() == 1:
            input_ids = input_ids[None]
        B, S = input_ids.shape
        if S > s.max_position_embeddings:
            raise ValueError(f"{S} tokens > max_position_embeddings {s.max_position_embeddings}")
        P = self._packed(dev)
        D, H = s.hidden_size, s.num_attention_heads
        d = D // H
        ids = input_ids.to(dev, torch.int32).reshape(-1).contiguous()
        h = be.embed_tokens(ids, P["tok"], P["pos"], S)
        hidden: List[torch.Tensor] = [h]
        kind = ACT_KINDS[s.hidden_act]
        for L in P["layers"]:
            y, _ = be.layer_norm(h, L["ln1"][0], L["ln1"][1], s.layer_norm_eps)
            qkv = be.gemm(y, L["wqkv"], bias=L["bqkv"])
            o, _ = be.attention_v0(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], B, S, S, H, d, d ** -0.5, causal=True)
            h = be.gemm(o, L["wo"], bias=L["bo"], residual=h)
            y, _ = be.layer_norm(h, L["ln2"][0], L["ln2"][1], s.layer_norm_eps)
            u = be.activation(be.gemm(y, L["w1"], bias=L["b1"]), kind)
            h = be.gemm(u, L["w2"], bias=L["b2"], residual=h)
            hidden.append(h)
        last, _ = be.layer_norm(h, P["lnf"][0], P["lnf"][1], s.layer_norm_eps)
        rows = self.eos_rows(input_ids).to(dev) + torch.arange(B, device=dev) * S
        pooled = last.index_select(0, rows).contiguous()
        text_embeds = be.gemm(pooled, P["proj"]) if P["proj"] is not None else None
        last3 = last.view(B, S, D)
        hs = tuple(t.view(B, S, D) for t in hidden) if output_hidden_states else None
        # CLIPTextModelOutput leads with text_embeds, BaseModelOutputWithPooling with last_hidden_state
        return _output(text_embeds if text_embeds is not None else last3, last3, pooled, text_embeds, hs)


def build_text_encoder(name: str, device="cuda", seed: int = 0, backend=None) -> ClipTextEncoder:
    """Seeded synthetic weights of a named architecture (no checkpoint exists offline): N(0, 0.02) embeddings and
    projections like CLIP's own init scale, LayerNorm gains near 1."""
    spec = TEXT_SPECS[name]
    enc = ClipTextEncoder(spec, backend=backend)
    g = torch.Generator().manual_seed(seed)
    for n, p in enc.named_parameters():
        if "layer_norm" in n:
            p.copy_((1.0 + 0.05 * torch.randn(p.shape, generator=g)) if n.endswith("weight")
                    else 0.02 * torch.randn(p.shape, generator=g))
        elif n.endswith("bias"):
            p.copy_(0.02 * torch.randn(p.shape, generator=g))
        elif "embedding" in n:
            p.copy_(0.02 * torch.randn(p.shape, generator=g))
        else:
            p.copy_(torch.randn(p.shape, generator=g) * (p.shape[1] ** -0.5))
    return enc.to(device)
