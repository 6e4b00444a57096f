"""Engine-backed UNet: the object handed to the reference as `unet`.

Drop-in boundary.  `EngineUNet` is what `lora.LoRANetwork(unet, ...)`
(lora.py:109-156) walks and what `train_util.predict_noise` (train_util.py:156-160) calls:

  * `named_modules()` yields a tree with diffusers' class names and attribute paths
    (`Transformer2DModel`, `ResnetBlock2D`, `Downsample2D`, `Upsample2D` holding real
    `nn.Linear` / `nn.Conv2d` children) so the reference's string matching (lora.py:62,68,
    188-197) finds the same 192 / 278 / 722 targets and exports the same key names;
  * the children only OWN the frozen weights.  Nothing here ever calls their `forward`:
    the network is executed by hand-written sm_100a kernels through `leco_b200.ops`
    (tcgen05 GEMM / implicit-GEMM conv with the LoRA residual as an extra K-segment,
    GroupNorm/LayerNorm/attention kernels);
  * `LoRAModule.apply_to` (lora.py:97-100) replaces `child.forward` by the adapter's bound
    method; the engine recognises that patch (`child.forward.__self__`) and reads
    `lora_down/lora_up/multiplier/scale` LIVE from it, so `with network:` (lora.py:231-237)
    and `optimizer.step()` behave exactly as with the reference UNet;
  * backward (only d(input) chains and the LoRA weight gradients exist — every other
    parameter is frozen, train_lora.py:69) is a reverse walk over a tape of engine ops.

The numerical backend is injected (`backend=`): the product uses `leco_b200.ops` (CUDA, no
fallback); the CPU test-suite injects a plain-torch double to validate this file's wiring
(forward order, skip connections, chain rule) against the oracle.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from types import SimpleNamespace
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn as nn

# --------------------------------------------------------------------------- configs
@dataclass
class UNetSpec:
    """Public unet/config.json values (diffusers 0.20) that define the topology."""
    name: str
    block_out_channels: Sequence[int] = (320, 640, 1280, 1280)
    attn_levels: Sequence[bool] = (True, True, True, False)
    layers_per_block: int = 2
    cross_attention_dim: int = 768
    num_heads: Sequence[int] = (8, 8, 8, 8)
    transformer_depth: Sequence[int] = (1, 1, 1, 1)
    use_linear_projection: bool = False
    norm_groups: int = 32
    text_time: bool = False  # SDXL addition_embed_type="text_time"
    add_time_dim: int = 256
    add_proj_in: int = 2816
    in_channels: int = 4
    out_channels: int = 4

    @property
    def temb_dim(self) -> int:
        return self.block_out_channels[0] * 4

    @property
    def add_text_dim(self) -> int:
        """Width of the pooled text embedding of `added_cond_kwargs["text_embeds"]` (SDXL: 2816 - 6*256 = 1280)."""
        return self.add_proj_in - 6 * self.add_time_dim


SPECS: Dict[str, UNetSpec] = {
    "sd15": UNetSpec("sd15", cross_attention_dim=768, num_heads=(8, 8, 8, 8)),
    "sd21": UNetSpec("sd21", cross_attention_dim=1024, num_heads=(5, 10, 20, 20), use_linear_projection=True),
    "sdxl": UNetSpec("sdxl", block_out_channels=(320, 640, 1280), attn_levels=(False, True, True),
                     cross_attention_dim=2048, num_heads=(5, 10, 20), transformer_depth=(1, 2, 10),
                     use_linear_projection=True, text_time=True),
    "tiny21": UNetSpec("tiny21", block_out_channels=(64, 128, 256, 256), cross_attention_dim=128,
                       num_heads=(1, 2, 4, 4), use_linear_projection=True),
    "tiny15": UNetSpec("tiny15", block_out_channels=(64, 128, 256, 256), cross_attention_dim=96,
                       num_heads=(8, 8, 8, 8)),
    "tinyxl": UNetSpec("tinyxl", block_out_channels=(64, 128, 256), attn_levels=(False, True, True),
                       cross_attention_dim=128, num_heads=(1, 2, 4), transformer_depth=(1, 2, 3),
                       use_linear_projection=True, text_time=True, add_time_dim=32, add_proj_in=32 * 6 + 64),
}


# --------------------------------------------------------------------------- module tree
# Parameter holders only.  Class names / attribute paths mirror diffusers because the
# reference matches on them (lora.py:188-197).  `forward` of a holder is never used.
class _Holder(nn.Module):
    def forward(self, *a, **k):  # pragma: no cover - engine never dispatches through modules
        raise RuntimeError("leco_b200 modules are parameter holders; call EngineUNet instead")


class TimestepEmbedding(_Holder):
    def __init__(self, cin, cout):
        super().__init__()
        self.linear_1 = nn.Linear(cin, cout)
        self.linear_2 = nn.Linear(cout, cout)


class ResnetBlock2D(_Holder):
    def __init__(self, cin, cout, temb, groups):
        super().__init__()
        self.norm1 = nn.GroupNorm(groups, cin, eps=1e-5)
        self.conv1 = nn.Conv2d(cin, cout, 3, 1, 1)
        self.time_emb_proj = nn.Linear(temb, cout)
        self.norm2 = nn.GroupNorm(groups, cout, eps=1e-5)
        self.conv2 = nn.Conv2d(cout, cout, 3, 1, 1)
        self.conv_shortcut = nn.Conv2d(cin, cout, 1, 1, 0) if cin != cout else None


class Downsample2D(_Holder):
    def __init__(self, ch):
        super().__init__()
        self.conv = nn.Conv2d(ch, ch, 3, 2, 1)


class Upsample2D(_Holder):
    def __init__(self, ch):
        super().__init__()
        self.conv = nn.Conv2d(ch, ch, 3, 1, 1)


class Attention(_Holder):
    def __init__(self, qdim, ctx_dim, heads, dim_head):
        super().__init__()
        inner = heads * dim_head
        self.heads, self.dim_head = heads, dim_head
        self.to_q = nn.Linear(qdim, inner, bias=False)
        self.to_k = nn.Linear(ctx_dim or qdim, inner, bias=False)
        self.to_v = nn.Linear(ctx_dim or qdim, inner, bias=False)
        self.to_out = nn.ModuleList([nn.Linear(inner, qdim), nn.Identity()])


class GEGLU(_Holder):
    def __init__(self, cin, cout):
        super().__init__()
        self.proj = nn.Linear(cin, cout * 2)


class FeedForward(_Holder):
    def __init__(self, dim):
        super().__init__()
        self.net = nn.ModuleList([GEGLU(dim, dim * 4), nn.Identity(), nn.Linear(dim * 4, dim)])


class BasicTransformerBlock(_Holder):
    def __init__(self, dim, heads, dim_head, ctx_dim):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim)
        self.attn1 = Attention(dim, None, heads, dim_head)
        self.norm2 = nn.LayerNorm(dim)
        self.attn2 = Attention(dim, ctx_dim, heads, dim_head)
        self.norm3 = nn.LayerNorm(dim)
        self.ff = FeedForward(dim)


class Transformer2DModel(_Holder):
    def __init__(self, heads, dim_head, ch, depth, ctx_dim, linear_proj, groups):
        super().__init__()
        inner = heads * dim_head
        self.norm = nn.GroupNorm(groups, ch, eps=1e-6)
        self.proj_in = nn.Linear(ch, inner) if linear_proj else nn.Conv2d(ch, inner, 1, 1, 0)
        self.transformer_blocks = nn.ModuleList(
            [BasicTransformerBlock(inner, heads, dim_head, ctx_dim) for _ in range(depth)])
        self.proj_out = nn.Linear(inner, ch) if linear_proj else nn.Conv2d(inner, ch, 1, 1, 0)


class _Block(_Holder):
    def __init__(self, resnets, attentions, sampler_name=None, sampler=None):
        super().__init__()
        self.resnets = nn.ModuleList(resnets)
        if attentions is not None:
            self.attentions = nn.ModuleList(attentions)
        if sampler is not None:
            setattr(self, sampler_name, nn.ModuleList([sampler]))


def _build_tree(root: nn.Module, s: UNetSpec):
    ch = list(s.block_out_channels)
    L = len(ch)
    temb = s.temb_dim

    def tr(c, heads, depth):
        return Transformer2DModel(heads, c // heads, c, depth, s.cross_attention_dim,
                                  s.use_linear_projection, s.norm_groups)

    root.conv_in = nn.Conv2d(s.in_channels, ch[0], 3, 1, 1)
    root.time_embedding = TimestepEmbedding(ch[0], temb)
    if s.text_time:
        root.add_embedding = TimestepEmbedding(s.add_proj_in, temb)
    downs = []
    out = ch[0]
    for i in range(L):
        cin, out = out, ch[i]
        res = [ResnetBlock2D(cin if j == 0 else out, out, temb, s.norm_groups) for j in range(s.layers_per_block)]
        att = [tr(out, s.num_heads[i], s.transformer_depth[i]) for _ in range(s.layers_per_block)] \
            if s.attn_levels[i] else None
        downs.append(_Block(res, att, "downsamplers", Downsample2D(out) if i != L - 1 else None))
    root.down_blocks = nn.ModuleList(downs)
    root.mid_block = _Block([ResnetBlock2D(ch[-1], ch[-1], temb, s.norm_groups) for _ in range(2)],
                            [tr(ch[-1], s.num_heads[-1], s.transformer_depth[-1])])
    ups = []
    rch, rheads = ch[::-1], list(s.num_heads)[::-1]
    rdepth, rattn = list(s.transformer_depth)[::-1], list(s.attn_levels)[::-1]
    out = rch[0]
    for i in range(L):
        prev, out = out, rch[i]
        cin = rch[min(i + 1, L - 1)]
        n = s.layers_per_block + 1
        res = []
        for j in range(n):
            skip = cin if j == n - 1 else out
            rin = prev if j == 0 else out
            res.append(ResnetBlock2D(rin + skip, out, temb, s.norm_groups))
        att = [tr(out, rheads[i], rdepth[i]) for _ in range(n)] if rattn[i] else None
        ups.append(_Block(res, att, "upsamplers", Upsample2D(out) if i != L - 1 else None))
    root.up_blocks = nn.ModuleList(ups)
    root.conv_norm_out = nn.GroupNorm(s.norm_groups, ch[0], eps=1e-5)
    root.conv_out = nn.Conv2d(ch[0], s.out_channels, 3, 1, 1)


# --------------------------------------------------------------------------- tape
class Act:
    """An activation tensor plus 'does anything trainable sit upstream of it'."""
    __slots__ = ("t", "rg")

    def __init__(self, t, rg=False):
        self.t, self.rg = t, rg


class Tape:
    """Reverse-mode tape over engine ops.  Gradients are keyed by tensor identity."""

    def __init__(self, be):
        self.be = be
        self.nodes: List[Callable[[], None]] = []
        self.grads: Dict[int, torch.Tensor] = {}
        self._keep: List[torch.Tensor] = []
        self._aliased: set = set()

    def record(self, fn):
        self.nodes.append(fn)

    def grad(self, t):
        return self.grads.get(id(t))

    def accum(self, t, g, owned: bool):
        """grads[t] += g.  `owned`=False means g is someone else's buffer (aliasing is allowed
        once per buffer; a second alias is copied so in-place accumulation stays safe)."""
        k = id(t)
        if k in self.grads:
            self.be.add_(self.grads[k], g)
            return
        if not owned:
            if id(g) in self._aliased:
                g = self.be.clone(g)
            else:
                self._aliased.add(id(g))
        self.grads[k] = g
        self._keep.append(t)

    def accum_cols(self, t, col0, g):
        """grads[t][:, col0:col0+g.shape[1]] = g   (disjoint column slices written once each)."""
        k = id(t)
        if k not in self.grads:
            self.grads[k] = self.be.zeros_like(t)
            self._keep.append(t)
        self.be.copy_cols(g, 0, self.grads[k], col0, g.shape[1])

    def backward(self):
        for fn in reversed(self.nodes):
            fn()
        self.nodes.clear()


# --------------------------------------------------------------------------- LoRA discovery
def find_adapter(child: nn.Module):
    """The LoRAModule whose bound `forward` replaced child.forward (lora.py:97-100), or None."""
    owner = getattr(getattr(child, "forward", None), "__self__", None)
    if owner is None or owner is child:
        return None
    if hasattr(owner, "lora_down") and hasattr(owner, "lora_up"):
        return owner
    return None


def down_as_rows(wd: torch.Tensor) -> torch.Tensor:
    """lora_down.weight -> [r, K] in the engine's k order (Linear: as is; Conv: (kh, kw, cin))."""
    if wd.dim() == 4:
        return wd.permute(0, 2, 3, 1).reshape(wd.shape[0], -1)
    return wd.reshape(wd.shape[0], -1)


def rows_as_down(rows: torch.Tensor, shape) -> torch.Tensor:
    """inverse of down_as_rows (a view when `rows` is contiguous)."""
    if len(shape) == 4:
        r, cin, kh, kw = shape
        return rows.reshape(r, kh, kw, cin).permute(0, 3, 1, 2)
    return rows.reshape(shape)


class LoraSite:
    """One fused GEMM site with up to three adapted nn.Linear / 1x1-conv members sharing an
    input (q,k,v / k,v) or a single member.  Holds the padded tensor-core operands:
        ad  [Kl, K]  rows = stacked lora_down weights (zero rows pad Kl to a multiple of 16)
        bup [N, Kl]  block structure: member i occupies rows n_off_i.., cols k_off_i..
    Values are re-copied from the adapter parameters whenever their version counters move."""

    def __init__(self, members: List[nn.Module], n_offsets: List[int], n_total: int, k_in: int):
        self.members, self.n_offsets, self.n_total, self.k_in = members, n_offsets, n_total, k_in
        # a 3x3 lora_down ([r, Cin, 3, 3], lora.py:76-81) is stored as [r, (kh, kw, Cin)]: the k order of
        # the implicit-GEMM conv
        self.ad = self.bup = None
        self.g_ad = self.g_bup = None
        self._versions = None
        self._native = False
        self._home_ptr = None
        self.ranks = None
        self.static_t = None   # (ad^T, bup^T) static views (leco_b200.lora.FlatState.build_transposed)
        self.flat_state = None

    def bind_native(self, ad, bup, g_ad, g_bup, home_ptr=None):
        """The adapter Parameters are views of `ad`/`bup` (leco_b200.lora flat layout): no packing,
        gradients accumulate straight into the flat fp32 buffer.  For a float32 network the Parameters are views of
        the fp32 master instead (`home_ptr` = where this site's first Parameter starts there) and `ad`/`bup` are the
        bf16 copies FlatState keeps current."""
        self.ad, self.bup, self.g_ad, self.g_bup = ad, bup, g_ad, g_bup
        self._home_ptr = ad.data_ptr() if home_ptr is None else home_ptr
        self.ranks = [a.lora_down.weight.shape[0] for a in self.adapters()]
        self._native = True

    def adapters(self):
        ads = [find_adapter(m) for m in self.members]
        if all(a is None for a in ads):
            return None
        if any(a is None for a in ads):
            raise RuntimeError("leco_b200: partially adapted fused projection group is unsupported")
        return ads

    def active(self):
        """(adapters, multiplier) if the LoRA branch contributes, else None (multiplier == 0 is
        skipped: exact)."""
        ads = self.adapters()
        if ads is None:
            return None
        mult = float(ads[0].multiplier)
        if any(float(a.multiplier) != mult for a in ads):
            raise RuntimeError("leco_b200: adapters of one fused site disagree on multiplier")
        if mult == 0.0:
            return None
        return ads, mult

    def refresh(self, ads, device, dtype):
        if self._native:
            if ads[0].lora_down.weight.data_ptr() == self._home_ptr:
                return self.ad, self.bup
            self._native = False  # parameters were re-homed (e.g. .to()): fall back to packing
            self.ad = self.bup = self.g_ad = self.g_bup = None
            self.static_t = self.flat_state = None
        ranks = [a.lora_down.weight.shape[0] for a in ads]
        kl = (sum(ranks) + 15) // 16 * 16
        vers = tuple((a.lora_down.weight._version, a.lora_up.weight._version, a.lora_down.weight.data_ptr())
                     for a in ads)
        if self.ad is None or self.ad.shape[0] != kl or self.ad.device != device:
            self.ad = torch.zeros((kl, self.k_in), device=device, dtype=dtype)
            self.bup = torch.zeros((self.n_total, kl), device=device, dtype=dtype)
            self._versions = None
        if vers != self._versions:
            k0 = 0
            for a, n0, r in zip(ads, self.n_offsets, ranks):
                wd = a.lora_down.weight.detach()
                wu = a.lora_up.weight.detach()
                self.ad[k0:k0 + r].copy_(down_as_rows(wd))
                n = wu.shape[0]
                self.bup[n0:n0 + n, k0:k0 + r].copy_(wu.reshape(n, r))
                k0 += r
            self._versions = vers
        self.ranks = ranks
        return self.ad, self.bup

    def transposed(self, be):
        """(ad^T [K, Kl], bup^T [Kl, N]) for the backward GEMMs.  A trainer that owns the parameters keeps them in
        static buffers refreshed ONCE per optimizer step (`static_t`, one launch for all sites: they only change
        there); otherwise they are formed on the fly."""
        if self.static_t is not None:
            return self.static_t
        return be.transpose2d(self.ad), be.transpose2d(self.bup)

    # fp32 gradient accumulators in operand layout (zeroed per backward)
    def begin_grad(self, be):
        if not self._native:
            self.g_ad = self.g_bup = None

    def grad_ad(self, be):
        if self.g_ad is None:
            self.g_ad = be.zeros(self.ad.shape, self.ad, torch.float32)
        return self.g_ad

    def grad_bup(self, be):
        if self.g_bup is None:
            self.g_bup = be.zeros(self.bup.shape, self.bup, torch.float32)
        return self.g_bup

    def collect_grads(self, be):
        """[d lora_down.weight, d lora_up.weight] per member, in the adapters' own shapes/dtypes."""
        out, k0 = [], 0
        ads = self.adapters()
        for a, n0, r in zip(ads, self.n_offsets, self.ranks):
            wd, wu = a.lora_down.weight, a.lora_up.weight
            n = wu.shape[0]
            gd = None if self.g_ad is None else rows_as_down(self.g_ad[k0:k0 + r], wd.shape).to(wd.dtype, copy=True)
            gu = None if self.g_bup is None else self.g_bup[n0:n0 + n, k0:k0 + r].reshape(wu.shape).to(wu.dtype, copy=True)
            out += [gd, gu]
            k0 += r
        if self._native:  # handed to autograd as copies: reset the flat accumulators
            self.g_ad.zero_()
            self.g_bup.zero_()
        else:
            self.g_ad = self.g_bup = None
        return out


# --------------------------------------------------------------------------- packed weights
class LinearPack:
    """Frozen weights of one GEMM site in kernel layout (bf16): w [N,K], wt [K,N], bias [N]."""

    def __init__(self, mods: List[nn.Module], device, dtype, need_wt=True):
        ws, bs = [], []
        for m in mods:
            w = m.weight.detach()
            ws.append(w.reshape(w.shape[0], -1))
            if m.bias is not None:
                bs.append(m.bias.detach())
        w = torch.cat(ws, 0).to(device=device, dtype=dtype).contiguous()
        self.w = w
        self.wt = w.t().contiguous() if need_wt else None
        self.bias = torch.cat(bs).to(device=device, dtype=dtype).contiguous() if bs else None
        n_off, o = [], 0
        for x in ws:
            n_off.append(o)
            o += x.shape[0]
        self.site = LoraSite(mods, n_off, o, w.shape[1])
        self.n, self.k = w.shape


class ConvPack:
    """3x3 conv in implicit-GEMM layout: w [O, 9*I] (k = tap*I + c); wt_flip [I, 9*O] for d(input)."""

    def __init__(self, m: nn.Conv2d, device, dtype):
        w = m.weight.detach().to(device=device, dtype=dtype)  # OIHW
        O, I = w.shape[0], w.shape[1]
        self.w = w.permute(0, 2, 3, 1).reshape(O, 9 * I).contiguous()
        # d(input)[p, i] = sum_{tap, o} dy[p - off(tap), o] w[o, i, tap] = conv3x3(dy, flipped/transposed w)
        self.wt_flip = w.flip(2, 3).permute(1, 2, 3, 0).reshape(I, 9 * O).contiguous()
        self.wt = self.w.t().contiguous()  # [9I, O] for the stride-2 (im2col) variant's d(col)
        self.bias = m.bias.detach().to(device=device, dtype=dtype).contiguous() if m.bias is not None else None
        self.stride = m.stride[0]
        self.cin, self.cout = I, O
        self.module = m
        self.site = LoraSite([m], [0], O, 9 * I)
        self.n, self.k = O, 9 * I


class NormPack:
    def __init__(self, m, device, dtype):
        self.gamma = m.weight.detach().to(device=device, dtype=dtype).contiguous()
        self.beta = m.bias.detach().to(device=device, dtype=dtype).contiguous()
        self.eps = m.eps
        self.groups = getattr(m, "num_groups", None)


TEMB_GROUP = 4


# --------------------------------------------------------------------------- the engine
class EngineUNet(nn.Module):
    """`unet` duck type of the reference: callable
    `unet(sample[2B,4,h,w], timestep, encoder_hidden_states=[2B,77,D], added_cond_kwargs=None).sample`,
    plus `.to / .requires_grad_ / .eval / .enable_xformers_memory_efficient_attention`."""

    def __init__(self, spec: UNetSpec, backend=None):
        super().__init__()
        self.spec = spec
        self.config = SimpleNamespace(in_channels=spec.in_channels,
                                      addition_embed_type="text_time" if spec.text_time else None)
        _build_tree(self, spec)
        self._be = backend
        self._packed = False
        self._act_dtype = torch.bfloat16
        self._last_tape: Optional[Tape] = None
        self.attention_impl = "auto"
        # LoRA branch computed inside the GEMM kernel (True: x.A^T rides along as extra accumulator columns, T.B^T is a second
        # UMMA on the same accumulator; measured 250 vs 264 ms / iteration) or as a separate T GEMM + extra K-segment (False)
        self.fused_lora = bool(int(__import__('os').environ.get('LECO_FUSED_LORA', '1')))

    # ---- reference-facing no-ops -------------------------------------------------------
    def enable_xformers_memory_efficient_attention(self, *a, **k):  # train_lora.py:68
        return None

    @property
    def backend(self):
        if self._be is None:
            from . import ops  # product path: CUDA only, raises if the library is missing
            self._be = ops
        return self._be

    # ---- weight packing ----------------------------------------------------------------
    def pack(self, device=None, dtype=None):
        """(Re)build kernel-layout copies of the frozen weights on `device`."""
        p = next(self.parameters())
        device = device or p.device
        dtype = dtype or self._act_dtype
        s = self.spec
        P = SimpleNamespace()
        P.conv_in_w = self.conv_in.weight.detach().to(device=device, dtype=dtype).contiguous()
        P.conv_in_b = self.conv_in.bias.detach().to(device=device, dtype=dtype).contiguous()
        P.time1 = LinearPack([self.time_embedding.linear_1], device, dtype, need_wt=False)
        P.time2 = LinearPack([self.time_embedding.linear_2], device, dtype, need_wt=False)
        if s.text_time:
            P.add1 = LinearPack([self.add_embedding.linear_1], device, dtype, need_wt=False)
            P.add2 = LinearPack([self.add_embedding.linear_2], device, dtype, need_wt=False)
        resnets: List[ResnetBlock2D] = [m for m in self.modules() if isinstance(m, ResnetBlock2D)]
        # all time_emb_proj layers share the input silu(emb): batched into GEMMs of TEMB_GROUP layers each
        # (a group's stacked adapter rank must fit one 64-wide K-segment: 4 layers x rank <= 16)
        P.temb = []
        P.res = {}
        for g0 in range(0, len(resnets), TEMB_GROUP):
            grp = resnets[g0:g0 + TEMB_GROUP]
            P.temb.append(LinearPack([r.time_emb_proj for r in grp], device, dtype, need_wt=False))
            off = 0
            for r in grp:
                P.res[id(r)] = SimpleNamespace(
                    norm1=NormPack(r.norm1, device, dtype), conv1=ConvPack(r.conv1, device, dtype),
                    norm2=NormPack(r.norm2, device, dtype), conv2=ConvPack(r.conv2, device, dtype),
                    shortcut=LinearPack([r.conv_shortcut], device, dtype) if r.conv_shortcut is not None else None,
                    temb_group=g0 // TEMB_GROUP, temb_off=off, cout=r.conv1.out_channels)
                off += r.conv1.out_channels
        P.tr = {}
        for t in (m for m in self.modules() if isinstance(m, Transformer2DModel)):
            tp = SimpleNamespace(norm=NormPack(t.norm, device, dtype),
                                 proj_in=LinearPack([t.proj_in], device, dtype),
                                 proj_out=LinearPack([t.proj_out], device, dtype), blocks=[])
            for b in t.transformer_blocks:
                tp.blocks.append(SimpleNamespace(
                    norm1=NormPack(b.norm1, device, dtype), norm2=NormPack(b.norm2, device, dtype),
                    norm3=NormPack(b.norm3, device, dtype),
                    qkv=LinearPack([b.attn1.to_q, b.attn1.to_k, b.attn1.to_v], device, dtype),
                    out1=LinearPack([b.attn1.to_out[0]], device, dtype),
                    q2=LinearPack([b.attn2.to_q], device, dtype),
                    kv2=LinearPack([b.attn2.to_k, b.attn2.to_v], device, dtype),
                    out2=LinearPack([b.attn2.to_out[0]], device, dtype),
                    ff1=LinearPack([b.ff.net[0].proj], device, dtype),
                    ff2=LinearPack([b.ff.net[2]], device, dtype),
                    heads=b.attn1.heads, dim_head=b.attn1.dim_head))
            P.tr[id(t)] = tp
        P.samp = {}
        for m in self.modules():
            if isinstance(m, (Downsample2D, Upsample2D)):
                P.samp[id(m)] = ConvPack(m.conv, device, dtype)
        P.norm_out = NormPack(self.conv_norm_out, device, dtype)
        wo = self.conv_out.weight.detach().to(device=device, dtype=dtype)
        P.conv_out_w = wo.permute(0, 2, 3, 1).reshape(wo.shape[0], 9, wo.shape[1]).contiguous()
        P.conv_out_b = self.conv_out.bias.detach().to(device=device, dtype=dtype).contiguous()
        old = getattr(self, "_P", None)
        self._P = P
        if old is not None:
            # a re-pack (weights edited / moved) keeps the LoraSite objects: an adapter network bound to them
            # (leco_b200.lora.bind_flat: flat operand views, static transposes) stays bound
            self._P = old
            old_owners = self._site_owners()
            self._P = P
            for o_new, o_old in zip(self._site_owners(), old_owners):
                o_new.site = o_old.site
        self._packed = True
        self._pack_device = device
        return self

    def _ensure_packed(self, device):
        if not self._packed or self._pack_device != device:
            self.pack(device)

    def _apply(self, fn, *a, **k):          # .to() / .cuda() / .half(): the kernel-layout copies are stale
        self._packed = False
        return super()._apply(fn, *a, **k)

    def load_state_dict(self, *a, **k):     # in-place weight edits after a pack
        self._packed = False
        return super().load_state_dict(*a, **k)

    # ---- ops with backward rules ---------------------------------------------------------
    def _tn_chunks(self, be, a, b, out, transpose_out=False):
        """out (+)= a^T b with b's columns (the stacked adapter rank) taken 64 at a time (tn_reduce's tile)."""
        kl = b.shape[1]
        for c0 in range(0, kl, 64):
            c1 = min(kl, c0 + 64)
            be.tn_reduce(a, b[:, c0:c1], out[c0:c1] if transpose_out else out[:, c0:c1], transpose_out=transpose_out)

    def _linear(self, be, tape, x: Act, pk: LinearPack, residual: Optional[Act] = None, geglu=False,
                conv_nhw=None, out=None):
        """y = x W^T + b (+ s*(x A^T) B^T) (+ residual)   — lora.py:102-106 fused.  The LoRA branch is an extra
        K-segment of the tensor-core GEMM when the site's stacked rank fits one (<= 64); larger ranks (the reference
        accepts any) run it as a second accumulate-GEMM on the result."""
        act = pk.site.active()
        kw = {}
        T = ad = bup = None
        sm = 1.0
        big = False
        rg_out = x.rg or (residual is not None and residual.rg) or act is not None
        need_grad = tape is not None and rg_out
        if act is not None:
            ads, mult = act
            ad, bup = pk.site.refresh(ads, x.t.device, x.t.dtype)
            sm = float(ads[0].scale) * mult
            big = ad.shape[0] > 64
            if self.fused_lora and not big:   # T = s*m x A^T is formed inside the GEMM kernel (extra accumulator columns)
                kw.update(fl_ad=ad, fl_bup=bup, fl_scale=sm, fl_rank=sum(pk.site.ranks))
                if need_grad:
                    T = be.empty((x.t.shape[0], ad.shape[0]), x.t)      # every row is written by the kernel (fl_t_out)
                    kw["fl_t_out"] = T
            else:
                T = be.gemm(x.t, ad, alpha=sm)                  # T = s*m * x A^T   [M, Kl]
                if not big:
                    kw.update(lora_t=T, lora_up=bup)
        res_t = None if residual is None else residual.t
        if big:
            pre = be.gemm(x.t, pk.w, bias=pk.bias, residual=res_t)
            if geglu:
                y = be.geglu_fwd(be.gemm(T, bup, residual=pre))
                if out is not None:
                    out.copy_(y)
                    y = out
            else:
                y = be.gemm(T, bup, out, residual=pre)
        else:
            y = be.gemm(x.t, pk.w, out, bias=pk.bias, residual=res_t, geglu=geglu, **kw)
        res = Act(y, need_grad)
        if need_grad:
            assert not geglu, "GEGLU epilogue is used on the no-grad path only"

            def bwd():
                gy = tape.grad(y)
                if gy is None:
                    return
                if residual is not None and residual.rg:
                    tape.accum(residual.t, gy, owned=False)
                kw2 = {}
                dT = adT = None
                if act is not None:
                    adT, bupT = pk.site.transposed(be)               # [K, Kl], [Kl, N]
                    if self.fused_lora and x.rg and not big:
                        # one kernel: dT = s*m dY B (saved) and dx = dY W + dT A
                        dT = be.empty((gy.shape[0], bupT.shape[0]), gy)
                        kw2.update(fl_ad=bupT, fl_bup=adT, fl_scale=sm, fl_rank=sum(pk.site.ranks), fl_t_out=dT)
                    else:
                        dT = be.gemm(gy, bupT, alpha=sm)             # s*m * dY B      [M, Kl]
                        if x.rg and not big:
                            kw2.update(lora_t=dT, lora_up=adT)       # + dT A
                    # dB[n,k] += sum_m dY[m,n] T[m,k]   (T already carries s*m)
                    self._tn_chunks(be, gy, T, pk.site.grad_bup(be))
                if x.rg:
                    dx = be.gemm(gy, pk.wt, **kw2)
                    if big:
                        dx = be.gemm(dT, adT, residual=dx)
                    tape.accum(x.t, dx, owned=True)
                if act is not None:
                    # dA[k,j] += sum_m dT[m,k] x[m,j]   -> accumulated transposed-in-place
                    self._tn_chunks(be, x.t, dT, pk.site.grad_ad(be), transpose_out=True)
            tape.record(bwd)
        return res

    def _conv3x3(self, be, tape, x: Act, pk: ConvPack, n, h, w, rowbias: Optional[Act] = None, rb_col0: int = 0,
                 residual: Optional[Act] = None):
        """3x3 / stride 1 / pad 1 conv as implicit GEMM (+ per-sample bias slice `rowbias[:, rb_col0:...]`,
        + residual).  A conv adapter (lora.py:68-82: 3x3 lora_down to r channels, 1x1 lora_up) adds
        T = s*m*conv3x3(x, A) as the extra K-segment (or, above rank 64, as a second accumulate-GEMM)."""
        act = pk.site.active()
        kw = {}
        T = ad = bup = None
        sm = 1.0
        big = False
        if act is not None:
            ads, mult = act
            ad, bup = pk.site.refresh(ads, x.t.device, x.t.dtype)
            sm = float(ads[0].scale) * mult
            big = ad.shape[0] > 64
            if self.fused_lora and not big:
                kw.update(fl_ad=ad, fl_bup=bup, fl_scale=sm, fl_rank=sum(pk.site.ranks))
                if tape is not None:
                    T = be.empty((x.t.shape[0], ad.shape[0]), x.t)      # every row is written by the kernel (fl_t_out)
                    kw["fl_t_out"] = T
            else:
                T = be.gemm(x.t, ad, alpha=sm, conv_nhw=(n, h, w))       # [M, Kl]
                if not big:
                    kw.update(lora_t=T, lora_up=bup)
        rb = None if rowbias is None else rowbias.t[:, rb_col0:rb_col0 + pk.cout]
        y = be.gemm(x.t, pk.w, bias=pk.bias, rowbias=rb, rows_per_group=h * w,
                    residual=None if residual is None else residual.t, conv_nhw=(n, h, w), **kw)
        if big:
            y = be.gemm(T, bup, residual=y)
        rg = x.rg or (residual is not None and residual.rg) or (rowbias is not None and rowbias.rg) or act is not None
        out = Act(y, rg and tape is not None)
        if out.rg:
            def bwd():
                gy = tape.grad(y)
                if gy is None:
                    return
                if residual is not None and residual.rg:
                    tape.accum(residual.t, gy, owned=False)
                if rowbias is not None and rowbias.rg:
                    tape.accum_cols(rowbias.t, rb_col0, be.rowgroup_sum(gy, n, h * w))
                dx_lora = None
                if act is not None:
                    _, bupT = pk.site.transposed(be)
                    dT = be.gemm(gy, bupT, alpha=sm)                             # s*m * dY B   [M, Kl]
                    self._tn_chunks(be, gy, T, pk.site.grad_bup(be))
                    self._tn_chunks(be, be.im2col_s1(x.t, n, h, w), dT, pk.site.grad_ad(be), transpose_out=True)
                    if x.rg:
                        kl = ad.shape[0]
                        # dx[p,c] += sum_{tap',kl} dT[p+off(tap'), kl] * A[kl, flip(tap'), c]
                        bmat = ad.reshape(kl, 9, pk.cin).flip(1).permute(2, 1, 0).reshape(pk.cin, 9 * kl).contiguous()
                        dx_lora = (be.im2col_s1(dT, n, h, w), bmat)
                if x.rg:
                    dx = be.gemm(gy, pk.wt_flip, conv_nhw=(n, h, w))
                    if dx_lora is not None:
                        dx = be.gemm(dx_lora[0], dx_lora[1], residual=dx)
                    tape.accum(x.t, dx, owned=True)
            tape.record(bwd)
        return out

    def _conv_s2(self, be, tape, x: Act, pk: ConvPack, n, h, w):
        """3x3 / stride 2 / pad 1 (Downsample2D): explicit im2col (3 small layers) + the linear site."""
        col_t = be.im2col_s2(x.t, n, h, w)
        col = Act(col_t, x.rg and tape is not None)
        out = self._linear(be, tape, col, pk)
        if tape is not None and x.rg:
            def bwd():
                g = tape.grad(col_t)
                if g is not None:
                    tape.accum(x.t, be.col2im_s2(g, n, h, w), owned=True)
            # recorded AFTER the linear node, so it runs BEFORE it in the reverse walk: re-order
            lin = tape.nodes.pop() if out.rg else None
            tape.record(bwd)
            if lin is not None:
                tape.record(lin)
        return out

    def _group_norm(self, be, tape, x: Act, pk: NormPack, n, hw, silu):
        y, stats = be.group_norm(x.t, n, hw, pk.gamma, pk.beta, pk.groups, pk.eps, silu)
        out = Act(y, x.rg and tape is not None)
        if out.rg:
            def bwd():
                gy = tape.grad(y)
                if gy is not None:
                    tape.accum(x.t, be.group_norm_bwd(x.t, gy, stats, pk.gamma, pk.beta, n, hw, pk.groups, silu),
                               owned=True)
            tape.record(bwd)
        return out

    def _layer_norm(self, be, tape, x: Act, pk: NormPack):
        want = x.rg and tape is not None
        y, stats = be.layer_norm(x.t, pk.gamma, pk.beta, pk.eps, want)
        out = Act(y, want)
        if want:
            def bwd():
                gy = tape.grad(y)
                if gy is not None:
                    tape.accum(x.t, be.layer_norm_bwd(x.t, gy, stats, pk.gamma), owned=True)
            tape.record(bwd)
        return out

    def _attention(self, be, tape, qa: Act, qc: int, ka: Act, kc: int, va: Act, vc: int, nb, sq, skv, heads, d):
        """softmax(Q K^T / sqrt(d)) V per (sample, head).  Q/K/V are column blocks (offsets qc/kc/vc,
        width heads*d) of the fused projection outputs owned by the Acts qa/ka/va."""
        C = heads * d
        scale = d ** -0.5
        qt, kt, vt = qa.t[:, qc:qc + C], ka.t[:, kc:kc + C], va.t[:, vc:vc + C]
        rg = (qa.rg or ka.rg or va.rg) and tape is not None
        o, saved = be.attention(qt, kt, vt, nb, sq, skv, heads, d, scale, rg)
        out = Act(o, rg)
        if rg:
            def bwd():
                go = tape.grad(o)
                if go is None:
                    return
                # the projections feed only this attention: their gradient buffers are written here
                bufs = {}
                for owner in (qa, ka, va):
                    if owner.rg and id(owner.t) not in bufs:
                        assert tape.grad(owner.t) is None
                        bufs[id(owner.t)] = be.empty_like(owner.t)

                def view(owner, c0):
                    return bufs[id(owner.t)][:, c0:c0 + C] if owner.rg else None
                be.attention_bwd(go, qt, kt, vt, saved, nb, sq, skv, heads, d, scale,
                                 view(qa, qc), view(ka, kc), view(va, vc))
                for owner in (qa, ka, va):
                    if owner.rg and id(owner.t) in bufs:
                        tape.accum(owner.t, bufs.pop(id(owner.t)), owned=True)
            tape.record(bwd)
        return out

    # ---- network pieces ---------------------------------------------------------------
    def _resnet(self, be, tape, x: Act, rp, temb_groups, n, h, w):
        hw = h * w
        hdn = self._group_norm(be, tape, x, rp.norm1, n, hw, True)
        hdn = self._conv3x3(be, tape, hdn, rp.conv1, n, h, w, rowbias=temb_groups[rp.temb_group], rb_col0=rp.temb_off)
        hdn = self._group_norm(be, tape, hdn, rp.norm2, n, hw, True)
        sc = x if rp.shortcut is None else self._linear(be, tape, x, rp.shortcut)
        return self._conv3x3(be, tape, hdn, rp.conv2, n, h, w, residual=sc)

    def _transformer(self, be, tape, x: Act, tp, ctx: Act, n, h, w, kv_cache=None):
        hw = h * w
        hdn = self._group_norm(be, tape, x, tp.norm, n, hw, False)
        hdn = self._linear(be, tape, hdn, tp.proj_in)
        for bp in tp.blocks:
            C = bp.heads * bp.dim_head
            n1 = self._layer_norm(be, tape, hdn, bp.norm1)
            qkv = self._linear(be, tape, n1, bp.qkv)
            a = self._attention(be, tape, qkv, 0, qkv, C, qkv, 2 * C, n, hw, hw, bp.heads, bp.dim_head)
            hdn = self._linear(be, tape, a, bp.out1, residual=hdn)
            n2 = self._layer_norm(be, tape, hdn, bp.norm2)
            q2 = self._linear(be, tape, n2, bp.q2)
            if kv_cache is not None:   # hoisted: K/V of the (per-iteration constant) text embedding, see cross_kv()
                kv = Act(kv_cache.pop(0), False)
            else:
                kv = self._linear(be, tape, ctx, bp.kv2)
            skv = ctx.t.shape[0] // n
            a = self._attention(be, tape, q2, 0, kv, 0, kv, C, n, hw, skv, bp.heads, bp.dim_head)
            hdn = self._linear(be, tape, a, bp.out2, residual=hdn)
            n3 = self._layer_norm(be, tape, hdn, bp.norm3)
            if tape is None:
                f = self._linear(be, None, n3, bp.ff1, geglu=True)
            else:
                pre = self._linear(be, tape, n3, bp.ff1)
                f = self._geglu(be, tape, pre)
            hdn = self._linear(be, tape, f, bp.ff2, residual=hdn)
        return self._linear(be, tape, hdn, tp.proj_out, residual=x)

    def _geglu(self, be, tape, pre: Act):
        y = be.geglu_fwd(pre.t)
        out = Act(y, pre.rg and tape is not None)
        if out.rg:
            def bwd():
                gy = tape.grad(y)
                if gy is not None:
                    tape.accum(pre.t, be.geglu_bwd(pre.t, gy), owned=True)
            tape.record(bwd)
        return out

    def _concat(self, be, tape, a: Act, b: Act):
        y = be.concat2(a.t, b.t)
        out = Act(y, (a.rg or b.rg) and tape is not None)
        if out.rg:
            ca = a.t.shape[1]

            def bwd():
                gy = tape.grad(y)
                if gy is None:
                    return
                ga, gb = be.split2(gy, ca)
                if a.rg:
                    tape.accum(a.t, ga, owned=True)
                if b.rg:
                    tape.accum(b.t, gb, owned=True)
            tape.record(bwd)
        return out

    def _upsample(self, be, tape, x: Act, n, h, w):
        y = be.upsample2x(x.t, n, h, w)
        out = Act(y, x.rg and tape is not None)
        if out.rg:
            def bwd():
                gy = tape.grad(y)
                if gy is not None:
                    tape.accum(x.t, be.upsample2x_bwd(gy, n, h, w), owned=True)
            tape.record(bwd)
        return out

    # ---- the forward program -----------------------------------------------------------
    def _transformers_in_order(self):
        """Transformer2DModel holders in forward-execution order (down, mid, up)."""
        out = []
        for blk in self.down_blocks:
            if hasattr(blk, "attentions"):
                out += list(blk.attentions)
        out += list(self.mid_block.attentions)
        for blk in self.up_blocks:
            if hasattr(blk, "attentions"):
                out += list(blk.attentions)
        return out

    def cross_kv(self, ctx2d: torch.Tensor, out: Optional[List[torch.Tensor]] = None) -> List[torch.Tensor]:
        """K|V projections (`attn2.to_k | to_v`, adapters included at their current multiplier) of the text embedding
        for every cross-attention layer in execution order.  They depend only on the prompt and the adapter state —
        not on the latents or the timestep — so a k-step denoise loop computes them once and
        hands the list to `run(kv_cache=...)`.  `out` = preallocated [rows, 2C] buffers to fill (CUDA-graph statics)."""
        self._ensure_packed(ctx2d.device)
        be, P = self.backend, self._P
        ctx = Act(ctx2d, False)
        res, i = [], 0
        for t in self._transformers_in_order():
            for bp in P.tr[id(t)].blocks:
                res.append(self._linear(be, None, ctx, bp.kv2, out=None if out is None else out[i]).t)
                i += 1
        return res

    def cross_kv_shapes(self, rows: int):
        return [(rows, 2 * bp.heads * bp.dim_head) for t in self._transformers_in_order() for bp in self._P.tr[id(t)].blocks]

    def run(self, sample: torch.Tensor, t: torch.Tensor, ctx2d: torch.Tensor, added: Optional[dict] = None,
            tape: Optional[Tape] = None, kv_cache: Optional[List[torch.Tensor]] = None):
        """sample NCHW [N,4,h,w] (fp32 or activation dtype), t fp32 [N], ctx2d [N*S, D] activation
        dtype -> (eps NCHW fp32 [N,4,h,w], final Act).  With `tape`, records the backward.  `kv_cache` = result of
        `cross_kv(ctx2d)` (no-grad passes only: the cross-attention adapters get no gradient through a cache)."""
        self._ensure_packed(sample.device)
        be, s, P = self.backend, self.spec, self._P
        N, _, H, W = sample.shape
        if kv_cache is not None:
            assert tape is None, "kv_cache is for no-grad passes"
            kv_cache = list(kv_cache)
        # -- time embedding (no trainable inputs in scope: time_emb_proj is only adapted by c3lier)
        te = be.timestep_embedding(t, s.block_out_channels[0])
        emb = be.gemm(be.silu(be.gemm(te, P.time1.w, bias=P.time1.bias)), P.time2.w, bias=P.time2.bias)
        if s.text_time:
            ids = added["time_ids"].reshape(-1).to(torch.float32)
            tid = be.timestep_embedding(ids, s.add_time_dim).reshape(N, -1)
            add_in = be.cat_cols(added["text_embeds"].to(tid.dtype), tid)
            aug = be.gemm(be.silu(be.gemm(add_in, P.add1.w, bias=P.add1.bias)), P.add2.w, bias=P.add2.bias)
            emb = be.add_(emb, aug)
        semb = Act(be.silu(emb), False)
        temb_all = [self._linear(be, tape, semb, pk) for pk in P.temb]   # per-group [N, sum Cout]; c3lier adapts these

        ctx = Act(ctx2d, False)
        x = Act(be.conv_in(sample, P.conv_in_w, P.conv_in_b), False)
        skips: List[Tuple[Act, int, int]] = [(x, H, W)]
        h, w = H, W
        for blk in self.down_blocks:
            for i, r in enumerate(blk.resnets):
                x = self._resnet(be, tape, x, P.res[id(r)], temb_all, N, h, w)
                if hasattr(blk, "attentions"):
                    x = self._transformer(be, tape, x, P.tr[id(blk.attentions[i])], ctx, N, h, w, kv_cache)
                skips.append((x, h, w))
            if hasattr(blk, "downsamplers"):
                x = self._conv_s2(be, tape, x, P.samp[id(blk.downsamplers[0])], N, h, w)
                h, w = h // 2, w // 2
                skips.append((x, h, w))
        mb = self.mid_block
        x = self._resnet(be, tape, x, P.res[id(mb.resnets[0])], temb_all, N, h, w)
        x = self._transformer(be, tape, x, P.tr[id(mb.attentions[0])], ctx, N, h, w, kv_cache)
        x = self._resnet(be, tape, x, P.res[id(mb.resnets[1])], temb_all, N, h, w)
        for blk in self.up_blocks:
            for i, r in enumerate(blk.resnets):
                sk, sh, sw = skips.pop()
                assert (sh, sw) == (h, w)
                x = self._concat(be, tape, x, sk)
                x = self._resnet(be, tape, x, P.res[id(r)], temb_all, N, h, w)
                if hasattr(blk, "attentions"):
                    x = self._transformer(be, tape, x, P.tr[id(blk.attentions[i])], ctx, N, h, w, kv_cache)
            if hasattr(blk, "upsamplers"):
                x = self._upsample(be, tape, x, N, h, w)
                h, w = h * 2, w * 2
                x = self._conv3x3(be, tape, x, P.samp[id(blk.upsamplers[0])], N, h, w)
        xn = self._group_norm(be, tape, x, P.norm_out, N, h * w, True)
        eps = be.conv_out(xn.t, P.conv_out_w, P.conv_out_b, N, h, w)
        if tape is not None and xn.rg:
            C0 = s.block_out_channels[0]

            def bwd():
                g = tape.grads.get("eps")
                if g is not None:
                    tape.accum(xn.t, be.conv_out_bwd(g, P.conv_out_w, C0), owned=True)
            tape.record(bwd)
        return eps

    # ---- reference call surface -----------------------------------------------------------
    def _site_owners(self):
        """The packs (LinearPack / ConvPack) that own a LoraSite, in a fixed order."""
        out = []
        P = self._P
        for tp in P.tr.values():
            out += [tp.proj_in, tp.proj_out]
            for bp in tp.blocks:
                out += [bp.qkv, bp.out1, bp.q2, bp.kv2, bp.out2, bp.ff1, bp.ff2]
        out += list(P.temb)
        for rp in P.res.values():
            out += [rp.conv1, rp.conv2]
            if rp.shortcut is not None:
                out.append(rp.shortcut)
        out += list(P.samp.values())
        return out

    def lora_sites(self) -> List[LoraSite]:
        """Every GEMM site an adapter can attach to (lierla: the transformer projections; c3lier adds the
        resnet convs / time_emb_proj / shortcuts and the down/up-sampler convs)."""
        return [o.site for o in self._site_owners()]

    def forward(self, sample, timestep, encoder_hidden_states=None, added_cond_kwargs=None, **_):
        """train_util.py:156-160 / :239-244.  Under autograd (grad enabled and an adapter with
        multiplier != 0 attached) the result carries a grad_fn that routes d(sample) into the
        adapters' `.grad` through the engine's own backward."""
        dev = sample.device
        self._ensure_packed(dev)
        be = self.backend
        N = sample.shape[0]
        t = timestep
        if not torch.is_tensor(t):
            t = torch.tensor([float(t)], dtype=torch.float32)
        t = t.reshape(-1).to(device=dev, dtype=torch.float32).expand(N).contiguous()
        ctx2d = encoder_hidden_states.to(device=dev, dtype=self._act_dtype).reshape(
            -1, encoder_hidden_states.shape[-1]).contiguous()
        x_in = sample.detach().contiguous()
        if x_in.dtype not in (torch.float32, self._act_dtype):
            x_in = x_in.float()
        sites = [s for s in self.lora_sites() if s.active() is not None]
        if torch.is_grad_enabled() and sites:
            for fs in {id(st.flat_state): st.flat_state for st in sites if st.flat_state is not None}.values():
                fs.refresh_transposed()
            params = []
            for st in sites:
                for a in st.adapters():
                    params += [a.lora_down.weight, a.lora_up.weight]
            eps = _EngineFn.apply(self, x_in, t, ctx2d, added_cond_kwargs, sites, *params)
        else:
            # a float32 network's operands are a bf16 COPY of its Parameters: bring it up to date (whoever stepped them)
            for fs in {id(st.flat_state): st.flat_state for st in sites if st.flat_state is not None}.values():
                fs.refresh_operands()
            eps = self.run(x_in, t, ctx2d, added_cond_kwargs, None)
        return SimpleNamespace(sample=eps.to(sample.dtype))


class _EngineFn(torch.autograd.Function):
    """One autograd node for the whole UNet: forward = engine program with a tape, backward =
    tape walk producing the adapters' weight gradients (the only trainable tensors)."""

    @staticmethod
    def forward(ctx, eng: EngineUNet, x, t, ctx2d, added, sites, *params):
        be = eng.backend
        tape = Tape(be)
        for st in sites:
            st.begin_grad(be)
        eps = eng.run(x, t, ctx2d, added, tape)
        ctx.tape, ctx.sites, ctx.be = tape, sites, be
        return eps

    @staticmethod
    def backward(ctx, g_eps):
        tape, be = ctx.tape, ctx.be
        tape.grads["eps"] = g_eps.contiguous().float()
        tape.backward()
        grads = []
        for st in ctx.sites:
            grads += st.collect_grads(be)
        return (None, None, None, None, None, None, *grads)
