"""ORACLE (test infrastructure, never on the product path).

Plain-PyTorch fp32 restatement of the third-party network the reference drives:
``diffusers==0.20.0`` ``UNet2DConditionModel`` (reference pin: requirements.txt:1;
call sites: train_util.py:156-160, train_util.py:239-244; loaded at
model_util.py:67-72, 169-174).  diffusers is NOT vendored under /root/reference
and is not installed in this image, so this file restates its published
algorithm from the public 0.20.0 sources (models/unet_2d_condition.py,
unet_2d_blocks.py, resnet.py, transformer_2d.py, attention.py,
attention_processor.py, embeddings.py).  PARITY UNPINNED for this third-party
part: no copy of diffusers exists here to check against; LECO's own arithmetic
(lora.py / train_util.py / prompt_util.py) IS pinned against the real reference
files through tests/golden (see tests/golden/make_golden.py).

The module tree reproduces diffusers' class names and attribute paths
(``Transformer2DModel``, ``ResnetBlock2D``, ``Downsample2D``, ``Upsample2D`` with
``Linear`` / ``Conv2d`` children named ``to_q``, ``to_out.0``, ``ff.net.0.proj`` ...)
because the reference discovers LoRA targets by class-name string
(lora.py:62,68,188,190) and derives the exported key names from the attribute
path (lora.py:191-192).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from types import SimpleNamespace
from typing import Optional, Sequence

import torch
import torch.nn as nn
import torch.nn.functional as F


# --------------------------------------------------------------------------- #
# configs (public unet/config.json values of SD1.5, SD2.1, SDXL-base)          #
# --------------------------------------------------------------------------- #
@dataclass
class UNetConfig:
    name: str
    in_channels: int = 4
    out_channels: int = 4
    block_out_channels: Sequence[int] = (320, 640, 1280, 1280)
    # True => CrossAttn{Down,Up}Block2D at that level, False => plain resnet block
    down_block_has_attn: Sequence[bool] = (True, True, True, False)
    layers_per_block: int = 2
    cross_attention_dim: int = 768
    # diffusers' "attention_head_dim" is really the HEAD COUNT per level
    num_heads: Sequence[int] = (8, 8, 8, 8)
    transformer_layers_per_block: Sequence[int] = (1, 1, 1, 1)
    use_linear_projection: bool = False
    norm_num_groups: int = 32
    addition_embed_type: Optional[str] = None  # "text_time" for SDXL
    addition_time_embed_dim: int = 256
    projection_class_embeddings_input_dim: int = 2816
    mid_block_layers: Optional[int] = None  # transformer depth of the mid block

    @property
    def time_embed_dim(self) -> int:
        return self.block_out_channels[0] * 4


CONFIGS = {
    "sd15": UNetConfig("sd15", cross_attention_dim=768, num_heads=(8, 8, 8, 8),
                       use_linear_projection=False),
    "sd21": UNetConfig("sd21", cross_attention_dim=1024, num_heads=(5, 10, 20, 20),
                       use_linear_projection=True),
    "sdxl": UNetConfig("sdxl", block_out_channels=(320, 640, 1280),
                       down_block_has_attn=(False, True, True),
                       cross_attention_dim=2048, num_heads=(5, 10, 20),
                       transformer_layers_per_block=(1, 2, 10),
                       use_linear_projection=True, addition_embed_type="text_time"),
    # reduced-width twins for fast CPU parity work: same topology, 64-wide heads
    "tiny21": UNetConfig("tiny21", block_out_channels=(64, 128, 256, 256),
                         cross_attention_dim=128, num_heads=(1, 2, 4, 4),
                         use_linear_projection=True),
    # SD1.x style: conv proj_in/out, 8 heads everywhere (head dims 8/16/32)
    "tiny15": UNetConfig("tiny15", block_out_channels=(64, 128, 256, 256),
                         cross_attention_dim=96, num_heads=(8, 8, 8, 8),
                         use_linear_projection=False),
    "tinyxl": UNetConfig("tinyxl", block_out_channels=(64, 128, 256),
                         down_block_has_attn=(False, True, True),
                         cross_attention_dim=128, num_heads=(1, 2, 4),
                         transformer_layers_per_block=(1, 2, 3),
                         use_linear_projection=True, addition_embed_type="text_time",
                         addition_time_embed_dim=32,
                         projection_class_embeddings_input_dim=32 * 6 + 64),
}


# --------------------------------------------------------------------------- #
# embeddings.py                                                                #
# --------------------------------------------------------------------------- #
def get_timestep_embedding(timesteps: torch.Tensor, dim: int, flip_sin_to_cos=True,
                           downscale_freq_shift=0.0, max_period=10000) -> torch.Tensor:
    """embeddings.py:get_timestep_embedding (flip_sin_to_cos=True, freq_shift=0 in SD)."""
    half = dim // 2
    exponent = -math.log(max_period) * torch.arange(half, dtype=torch.float32,
                                                    device=timesteps.device)
    exponent = exponent / (half - downscale_freq_shift)
    emb = timesteps[:, None].float() * torch.exp(exponent)[None, :]
    emb = torch.cat([torch.sin(emb), torch.cos(emb)], dim=-1)
    if flip_sin_to_cos:
        emb = torch.cat([emb[:, half:], emb[:, :half]], dim=-1)
    return emb


class Timesteps(nn.Module):
    def __init__(self, num_channels: int):
        super().__init__()
        self.num_channels = num_channels

    def forward(self, timesteps):
        return get_timestep_embedding(timesteps, self.num_channels)


class TimestepEmbedding(nn.Module):
    def __init__(self, in_channels: int, time_embed_dim: int):
        super().__init__()
        self.linear_1 = nn.Linear(in_channels, time_embed_dim)
        self.act = nn.SiLU()
        self.linear_2 = nn.Linear(time_embed_dim, time_embed_dim)

    def forward(self, sample):
        return self.linear_2(self.act(self.linear_1(sample)))


# --------------------------------------------------------------------------- #
# resnet.py                                                                    #
# --------------------------------------------------------------------------- #
class ResnetBlock2D(nn.Module):
    """GroupNorm(eps 1e-5) -> SiLU -> conv3x3 -> +time_emb_proj(SiLU(temb)) ->
    GroupNorm -> SiLU -> conv3x3, plus 1x1 conv_shortcut when channels change."""

    def __init__(self, in_channels: int, out_channels: int, temb_channels: int, groups=32,
                 eps=1e-5):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.norm1 = nn.GroupNorm(groups, in_channels, eps=eps, affine=True)
        self.conv1 = nn.Conv2d(in_channels, out_channels, 3, 1, 1)
        self.time_emb_proj = nn.Linear(temb_channels, out_channels)
        self.norm2 = nn.GroupNorm(groups, out_channels, eps=eps, affine=True)
        self.dropout = nn.Dropout(0.0)
        self.conv2 = nn.Conv2d(out_channels, out_channels, 3, 1, 1)
        self.nonlinearity = nn.SiLU()
        self.conv_shortcut = None
        if in_channels != out_channels:
            self.conv_shortcut = nn.Conv2d(in_channels, out_channels, 1, 1, 0)

    def forward(self, input_tensor, temb):
        h = self.conv1(self.nonlinearity(self.norm1(input_tensor)))
        t = self.time_emb_proj(self.nonlinearity(temb))[:, :, None, None]
        h = h + t
        h = self.conv2(self.dropout(self.nonlinearity(self.norm2(h))))
        if self.conv_shortcut is not None:
            input_tensor = self.conv_shortcut(input_tensor)
        return input_tensor + h  # output_scale_factor == 1.0


class Downsample2D(nn.Module):
    def __init__(self, channels: int):
        super().__init__()
        self.conv = nn.Conv2d(channels, channels, 3, 2, 1)

    def forward(self, x):
        return self.conv(x)


class Upsample2D(nn.Module):
    def __init__(self, channels: int):
        super().__init__()
        self.conv = nn.Conv2d(channels, channels, 3, 1, 1)

    def forward(self, x):
        return self.conv(F.interpolate(x, scale_factor=2.0, mode="nearest"))


# --------------------------------------------------------------------------- #
# attention_processor.py / attention.py                                        #
# --------------------------------------------------------------------------- #
class Attention(nn.Module):
    """to_q/to_k/to_v without bias, to_out.0 with bias; softmax(QK^T / sqrt(d)) V.
    The reference always enables xformers (train_lora.py:68): same maths."""

    def __init__(self, query_dim: int, cross_attention_dim: Optional[int], heads: int,
                 dim_head: int):
        super().__init__()
        inner = heads * dim_head
        ctx = cross_attention_dim if cross_attention_dim is not None else query_dim
        self.heads, self.dim_head = heads, dim_head
        self.scale = dim_head ** -0.5
        self.to_q = nn.Linear(query_dim, inner, bias=False)
        self.to_k = nn.Linear(ctx, inner, bias=False)
        self.to_v = nn.Linear(ctx, inner, bias=False)
        self.to_out = nn.ModuleList([nn.Linear(inner, query_dim), nn.Dropout(0.0)])

    def forward(self, hidden_states, encoder_hidden_states=None):
        ctx = hidden_states if encoder_hidden_states is None else encoder_hidden_states
        b, s, _ = hidden_states.shape
        q = self.to_q(hidden_states)
        k = self.to_k(ctx)
        v = self.to_v(ctx)

        def split(t):
            return t.reshape(b, t.shape[1], self.heads, self.dim_head).transpose(1, 2)

        q, k, v = split(q), split(k), split(v)
        attn = torch.softmax((q @ k.transpose(-1, -2)) * self.scale, dim=-1)
        o = (attn @ v).transpose(1, 2).reshape(b, s, self.heads * self.dim_head)
        return self.to_out[1](self.to_out[0](o))


class GEGLU(nn.Module):
    def __init__(self, dim_in: int, dim_out: int):
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out * 2)

    def forward(self, x):
        h, gate = self.proj(x).chunk(2, dim=-1)
        return h * F.gelu(gate)  # exact erf GELU


class FeedForward(nn.Module):
    def __init__(self, dim: int, mult: int = 4):
        super().__init__()
        inner = dim * mult
        self.net = nn.ModuleList([GEGLU(dim, inner), nn.Dropout(0.0), nn.Linear(inner, dim)])

    def forward(self, x):
        for m in self.net:
            x = m(x)
        return x


class BasicTransformerBlock(nn.Module):
    def __init__(self, dim: int, heads: int, dim_head: int, cross_attention_dim: int):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim)
        self.attn1 = Attention(dim, None, heads, dim_head)
        self.norm2 = nn.LayerNorm(dim)
        self.attn2 = Attention(dim, cross_attention_dim, heads, dim_head)
        self.norm3 = nn.LayerNorm(dim)
        self.ff = FeedForward(dim)

    def forward(self, h, encoder_hidden_states):
        h = self.attn1(self.norm1(h)) + h
        h = self.attn2(self.norm2(h), encoder_hidden_states) + h
        h = self.ff(self.norm3(h)) + h
        return h


class Transformer2DModel(nn.Module):
    """GroupNorm(eps 1e-6) -> proj_in -> blocks -> proj_out -> + residual.
    use_linear_projection=False (SD1.x): 1x1 Conv2d proj applied in NCHW before the
    token reshape; True (SD2.x/XL): Linear applied after it."""

    def __init__(self, heads: int, dim_head: int, in_channels: int, num_layers: int,
                 cross_attention_dim: int, use_linear_projection: bool, groups=32):
        super().__init__()
        inner = heads * dim_head
        self.use_linear_projection = use_linear_projection
        self.norm = nn.GroupNorm(groups, in_channels, eps=1e-6, affine=True)
        if use_linear_projection:
            self.proj_in = nn.Linear(in_channels, inner)
        else:
            self.proj_in = nn.Conv2d(in_channels, inner, 1, 1, 0)
        self.transformer_blocks = nn.ModuleList(
            [BasicTransformerBlock(inner, heads, dim_head, cross_attention_dim)
             for _ in range(num_layers)])
        if use_linear_projection:
            self.proj_out = nn.Linear(inner, in_channels)
        else:
            self.proj_out = nn.Conv2d(inner, in_channels, 1, 1, 0)

    def forward(self, hidden_states, encoder_hidden_states):
        b, c, hh, ww = hidden_states.shape
        residual = hidden_states
        h = self.norm(hidden_states)
        if not self.use_linear_projection:
            h = self.proj_in(h)
            inner = h.shape[1]
            h = h.permute(0, 2, 3, 1).reshape(b, hh * ww, inner)
        else:
            inner = c
            h = h.permute(0, 2, 3, 1).reshape(b, hh * ww, inner)
            h = self.proj_in(h)
        for blk in self.transformer_blocks:
            h = blk(h, encoder_hidden_states)
        if not self.use_linear_projection:
            h = h.reshape(b, hh, ww, inner).permute(0, 3, 1, 2).contiguous()
            h = self.proj_out(h)
        else:
            h = self.proj_out(h)
            h = h.reshape(b, hh, ww, inner).permute(0, 3, 1, 2).contiguous()
        return h + residual


# --------------------------------------------------------------------------- #
# unet_2d_blocks.py                                                            #
# --------------------------------------------------------------------------- #
class DownBlock(nn.Module):
    """CrossAttnDownBlock2D (has_attn) / DownBlock2D."""

    def __init__(self, cfg: UNetConfig, in_ch: int, out_ch: int, heads: int, depth: int,
                 has_attn: bool, add_downsample: bool):
        super().__init__()
        self.has_attn = has_attn
        self.resnets = nn.ModuleList()
        if has_attn:
            self.attentions = nn.ModuleList()
        for i in range(cfg.layers_per_block):
            self.resnets.append(ResnetBlock2D(in_ch if i == 0 else out_ch, out_ch,
                                              cfg.time_embed_dim, cfg.norm_num_groups))
            if has_attn:
                self.attentions.append(Transformer2DModel(
                    heads, out_ch // heads, out_ch, depth, cfg.cross_attention_dim,
                    cfg.use_linear_projection, cfg.norm_num_groups))
        self.downsamplers = None
        if add_downsample:
            self.downsamplers = nn.ModuleList([Downsample2D(out_ch)])

    def forward(self, h, temb, ctx):
        outs = []
        for i, resnet in enumerate(self.resnets):
            h = resnet(h, temb)
            if self.has_attn:
                h = self.attentions[i](h, ctx)
            outs.append(h)
        if self.downsamplers is not None:
            h = self.downsamplers[0](h)
            outs.append(h)
        return h, outs


class MidBlock(nn.Module):
    """UNetMidBlock2DCrossAttn: resnet, attention, resnet."""

    def __init__(self, cfg: UNetConfig, ch: int, heads: int, depth: int):
        super().__init__()
        self.resnets = nn.ModuleList([
            ResnetBlock2D(ch, ch, cfg.time_embed_dim, cfg.norm_num_groups),
            ResnetBlock2D(ch, ch, cfg.time_embed_dim, cfg.norm_num_groups)])
        self.attentions = nn.ModuleList([Transformer2DModel(
            heads, ch // heads, ch, depth, cfg.cross_attention_dim,
            cfg.use_linear_projection, cfg.norm_num_groups)])

    def forward(self, h, temb, ctx):
        h = self.resnets[0](h, temb)
        h = self.attentions[0](h, ctx)
        return self.resnets[1](h, temb)


class UpBlock(nn.Module):
    """CrossAttnUpBlock2D (has_attn) / UpBlock2D; skip tensors are concatenated
    AFTER the running hidden state: cat([hidden, skip], dim=1)."""

    def __init__(self, cfg: UNetConfig, in_ch: int, out_ch: int, prev_out_ch: int, heads: int,
                 depth: int, has_attn: bool, add_upsample: bool):
        super().__init__()
        self.has_attn = has_attn
        n = cfg.layers_per_block + 1
        self.resnets = nn.ModuleList()
        if has_attn:
            self.attentions = nn.ModuleList()
        for i in range(n):
            skip_ch = in_ch if i == n - 1 else out_ch
            res_in = prev_out_ch if i == 0 else out_ch
            self.resnets.append(ResnetBlock2D(res_in + skip_ch, out_ch, cfg.time_embed_dim,
                                              cfg.norm_num_groups))
            if has_attn:
                self.attentions.append(Transformer2DModel(
                    heads, out_ch // heads, out_ch, depth, cfg.cross_attention_dim,
                    cfg.use_linear_projection, cfg.norm_num_groups))
        self.upsamplers = None
        if add_upsample:
            self.upsamplers = nn.ModuleList([Upsample2D(out_ch)])

    def forward(self, h, skips, temb, ctx):
        for i, resnet in enumerate(self.resnets):
            h = torch.cat([h, skips.pop()], dim=1)
            h = resnet(h, temb)
            if self.has_attn:
                h = self.attentions[i](h, ctx)
        if self.upsamplers is not None:
            h = self.upsamplers[0](h)
        return h


# --------------------------------------------------------------------------- #
# unet_2d_condition.py                                                         #
# --------------------------------------------------------------------------- #
class UNet2DConditionModel(nn.Module):
    def __init__(self, cfg: UNetConfig):
        super().__init__()
        self.cfg = cfg
        self.config = SimpleNamespace(in_channels=cfg.in_channels,
                                      addition_embed_type=cfg.addition_embed_type)
        ch = list(cfg.block_out_channels)
        n_levels = len(ch)
        self.conv_in = nn.Conv2d(cfg.in_channels, ch[0], 3, 1, 1)
        self.time_proj = Timesteps(ch[0])
        self.time_embedding = TimestepEmbedding(ch[0], cfg.time_embed_dim)
        if cfg.addition_embed_type == "text_time":
            self.add_time_proj = Timesteps(cfg.addition_time_embed_dim)
            self.add_embedding = TimestepEmbedding(cfg.projection_class_embeddings_input_dim,
                                                   cfg.time_embed_dim)
        self.down_blocks = nn.ModuleList()
        out_ch = ch[0]
        for i in range(n_levels):
            in_ch, out_ch = out_ch, ch[i]
            self.down_blocks.append(DownBlock(
                cfg, in_ch, out_ch, cfg.num_heads[i], cfg.transformer_layers_per_block[i],
                cfg.down_block_has_attn[i], add_downsample=(i != n_levels - 1)))
        mid_depth = cfg.mid_block_layers or cfg.transformer_layers_per_block[-1]
        self.mid_block = MidBlock(cfg, ch[-1], cfg.num_heads[-1], mid_depth)
        self.up_blocks = nn.ModuleList()
        rev_ch = ch[::-1]
        rev_heads = list(cfg.num_heads)[::-1]
        rev_depth = list(cfg.transformer_layers_per_block)[::-1]
        rev_attn = list(cfg.down_block_has_attn)[::-1]
        out_ch = rev_ch[0]
        for i in range(n_levels):
            prev_out = out_ch
            out_ch = rev_ch[i]
            in_ch = rev_ch[min(i + 1, n_levels - 1)]
            self.up_blocks.append(UpBlock(
                cfg, in_ch, out_ch, prev_out, rev_heads[i], rev_depth[i], rev_attn[i],
                add_upsample=(i != n_levels - 1)))
        self.conv_norm_out = nn.GroupNorm(cfg.norm_num_groups, ch[0], eps=1e-5)
        self.conv_act = nn.SiLU()
        self.conv_out = nn.Conv2d(ch[0], cfg.out_channels, 3, 1, 1)

    # the reference calls these unconditionally (train_lora.py:67-70)
    def enable_xformers_memory_efficient_attention(self, *a, **k):
        return None

    def time_embed(self, sample, timestep, added_cond_kwargs=None):
        t = timestep
        if not torch.is_tensor(t):
            t = torch.tensor([t], dtype=torch.float32, device=sample.device)
        t = t.reshape(-1).to(sample.device).expand(sample.shape[0])
        emb = self.time_embedding(self.time_proj(t).to(sample.dtype))
        if self.cfg.addition_embed_type == "text_time":
            text_embeds = added_cond_kwargs["text_embeds"]
            time_ids = added_cond_kwargs["time_ids"]
            te = self.add_time_proj(time_ids.flatten()).reshape(text_embeds.shape[0], -1)
            add = torch.cat([text_embeds, te.to(text_embeds.dtype)], dim=-1).to(emb.dtype)
            emb = emb + self.add_embedding(add)
        return emb

    def forward(self, sample, timestep, encoder_hidden_states, added_cond_kwargs=None):
        emb = self.time_embed(sample, timestep, added_cond_kwargs)
        h = self.conv_in(sample)
        skips = [h]
        for blk in self.down_blocks:
            h, outs = blk(h, emb, encoder_hidden_states)
            skips.extend(outs)
        h = self.mid_block(h, emb, encoder_hidden_states)
        for blk in self.up_blocks:
            h = blk(h, skips, emb, encoder_hidden_states)
        h = self.conv_out(self.conv_act(self.conv_norm_out(h)))
        return SimpleNamespace(sample=h)


# --------------------------------------------------------------------------- #
# synthetic weights (no checkpoints exist in this sandbox)                     #
# --------------------------------------------------------------------------- #
def init_synthetic(model: nn.Module, seed: int = 0) -> nn.Module:
    """Seeded synthetic weights that keep activations O(1) and exercise every
    parameter: weights ~ N(0, g/fan_in), biases ~ N(0, 0.05^2), norm gamma ~
    1+0.1N, beta ~ 0.1N.  Residual-branch output layers are damped so the depth
    of the network does not blow the variance up.  Deterministic for a given
    torch build (CPU generator)."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    damp = ("conv2", "proj_out", "to_out.0", "ff.net.2")
    with torch.no_grad():
        for name, mod in model.named_modules():
            if isinstance(mod, (nn.Linear, nn.Conv2d)):
                w = mod.weight
                fan_in = w[0].numel()
                gain = 0.35 if name.endswith(damp) else 1.0
                w.copy_(torch.randn(w.shape, generator=g) * (gain / math.sqrt(fan_in)))
                if mod.bias is not None:
                    mod.bias.copy_(torch.randn(mod.bias.shape, generator=g) * 0.05)
            elif isinstance(mod, (nn.GroupNorm, nn.LayerNorm)):
                mod.weight.copy_(1.0 + 0.1 * torch.randn(mod.weight.shape, generator=g))
                mod.bias.copy_(0.1 * torch.randn(mod.bias.shape, generator=g))
    return model


def build_unet(cfg_name: str, seed: int = 0) -> UNet2DConditionModel:
    torch_state = torch.get_rng_state()
    model = UNet2DConditionModel(CONFIGS[cfg_name])
    torch.set_rng_state(torch_state)  # construction must not disturb the caller's RNG stream
    init_synthetic(model, seed)
    model.requires_grad_(False)
    model.eval()
    return model
