"""LoRA adapter network with the reference's call surface, stored the way the kernels want it.

Host-side mirror of lora.py (reference) for the drop-in boundary:

    LoRANetwork(unet, rank=4, multiplier=1.0, alpha=1.0, train_method="full")   lora.py:110-117
    .unet_loras[i].{lora_name, lora_dim, lora_down, lora_up, alpha, scale, multiplier, org_forward}
    .prepare_optimizer_params() -> [{"params": [...]}]                          lora.py:201-210
    .save_weights(file, dtype=None, metadata=None)                               lora.py:212-229
    with network: ...   (multiplier 1.0 inside, 0 outside)                       lora.py:231-237

Same discovery rule (class-name strings, outer-block train_method filter, lora.py:169-197),
same key names / order, same initialisation draws from the global CPU generator
(kaiming-uniform(a=sqrt 5) down, zeros up, lora.py:91-92), same error behaviour.

What is different is the storage.  When the network sits on a CUDA device and adapts a
`leco_b200.unet.EngineUNet`, all adapter weights live in ONE flat bf16 buffer laid out as
the tensor-core operands of the engine's fused GEMM sites (`ad [Kl,K]`, `bup [N,Kl]` per
site, see unet.LoraSite); every `lora_down.weight` / `lora_up.weight` Parameter is a VIEW
into it.  Consequences: the forward needs no packing, the backward accumulates straight
into a flat fp32 gradient buffer, AdamW is one fused kernel over the flat buffer
(`FlatAdamW`) and data-parallel training all-reduces that one buffer.

`network.to(device, dtype=torch.float32)` (`train.precision: float32`, the notebook's setting) keeps the Parameters as
views of a flat fp32 MASTER buffer instead; the bf16 operand buffer is then a copy the optimizer kernel rewrites with
every step (`leco_optim_flat_master`) and `FlatState.refresh_transposed` re-derives after any other in-place edit.
The tensor cores still compute on bf16 operands with fp32 accumulation — fp32 here is the precision of the adapter
weights, their gradients and the optimizer state, not of the UNet arithmetic.
"""
from __future__ import annotations

import math
import os
from typing import List, Optional

import torch
import torch.nn as nn

from .unet import down_as_rows, rows_as_down

PREFIX = "lora_unet"
LORA_PREFIX_UNET = PREFIX   # the reference's name (lora.py:16)
MAX_SITE_RANK = 64   # stacked (padded) adapter rank that rides as ONE tensor-core K-segment; larger ranks work too
                     # (unet._linear / _conv3x3 run the LoRA branch as a second accumulate-GEMM then)
# The reference keeps these as module-level lists and EXTENDS THE FIRST IN PLACE for c3lier
# (train_lora.py:44-46): callers that do the same to `DEFAULT_TARGET_REPLACE` here get the
# same behaviour, because the constructor reads the list at call time.
UNET_TARGET_REPLACE_MODULE_TRANSFORMER = ["Transformer2DModel"]
UNET_TARGET_REPLACE_MODULE_CONV = ["ResnetBlock2D", "Downsample2D", "Upsample2D"]
DEFAULT_TARGET_REPLACE = UNET_TARGET_REPLACE_MODULE_TRANSFORMER
TRAINING_METHODS = ("noxattn", "innoxattn", "selfattn", "xattn", "full")


def _skip_block(method: str, block_name: str) -> bool:
    if method == "full":
        return False
    if method == "noxattn":
        return "attn2" in block_name or "time_embed" in block_name
    if method == "innoxattn":
        return "attn2" in block_name
    if method == "selfattn":
        return "attn1" not in block_name
    if method == "xattn":
        return "attn2" not in block_name
    raise NotImplementedError(f"train_method: {method} is not implemented.")


class LoRAModule(nn.Module):
    """One adapter: y = org(x) + up(down(x)) * multiplier * scale, scale = alpha / rank."""

    def __init__(self, lora_name: str, org_module: nn.Module, multiplier=1.0, lora_dim=4, alpha=1):
        super().__init__()
        self.lora_name = lora_name
        self.lora_dim = lora_dim
        kind = org_module.__class__.__name__
        if kind == "Linear":
            self.lora_down = nn.Linear(org_module.in_features, lora_dim, bias=False)
            self.lora_up = nn.Linear(lora_dim, org_module.out_features, bias=False)
        elif kind == "Conv2d":
            cin, cout = org_module.in_channels, org_module.out_channels
            self.lora_dim = min(lora_dim, cin, cout)
            if self.lora_dim != lora_dim:
                print(f"{lora_name} dim (rank) is changed to: {self.lora_dim}")
            self.lora_down = nn.Conv2d(cin, self.lora_dim, org_module.kernel_size, org_module.stride,
                                       org_module.padding, bias=False)
            self.lora_up = nn.Conv2d(self.lora_dim, cout, (1, 1), (1, 1), bias=False)
        else:
            raise TypeError(f"cannot adapt a {kind}")
        if isinstance(alpha, torch.Tensor):
            alpha = alpha.detach().numpy()
        alpha = lora_dim if (alpha is None or alpha == 0) else alpha
        self.scale = alpha / self.lora_dim
        self.register_buffer("alpha", torch.tensor(alpha))
        nn.init.kaiming_uniform_(self.lora_down.weight, a=math.sqrt(5))
        nn.init.zeros_(self.lora_up.weight)
        self.multiplier = multiplier
        self._pending_org = [org_module]

    def apply_to(self):
        org = self._pending_org.pop()
        self.org_forward = org.forward
        org.forward = self.forward  # the engine recognises this patch (unet.find_adapter)

    def forward(self, x):
        # only reached when the adapted module is a real torch layer (not an engine holder)
        return self.org_forward(x) + self.lora_up(self.lora_down(x)) * self.multiplier * self.scale


class LoRANetwork(nn.Module):
    def __init__(self, unet, rank: int = 4, multiplier: float = 1.0, alpha: float = 1.0,
                 train_method: str = "full") -> None:
        super().__init__()
        self.multiplier, self.lora_dim, self.alpha = multiplier, rank, alpha
        targets = list(DEFAULT_TARGET_REPLACE)  # read at call time (c3lier aliasing)
        self.unet_loras: List[LoRAModule] = []
        for name, module in unet.named_modules():
            if _skip_block(train_method, name):
                continue
            if module.__class__.__name__ not in targets:
                continue
            for child_name, child in module.named_modules():
                if child.__class__.__name__ in ("Linear", "Conv2d"):
                    lora_name = f"{PREFIX}.{name}.{child_name}".replace(".", "_")
                    print(lora_name)
                    self.unet_loras.append(LoRAModule(lora_name, child, multiplier, rank, alpha))
        print(f"create LoRA for U-Net: {len(self.unet_loras)} modules.")
        seen = set()
        for lora in self.unet_loras:
            assert lora.lora_name not in seen, f"duplicated lora name: {lora.lora_name}. {seen}"
            seen.add(lora.lora_name)
        for lora in self.unet_loras:
            lora.apply_to()
            self.add_module(lora.lora_name, lora)
        self._engine = [unet] if hasattr(unet, "lora_sites") else []
        self.flat = None  # FlatState once bound

    # ---- reference surface ----------------------------------------------------------------
    def prepare_optimizer_params(self):
        if not self.unet_loras:
            return []
        return [{"params": [p for lora in self.unet_loras for p in lora.parameters()]}]

    def save_weights(self, file, dtype=None, metadata: Optional[dict] = None):
        sd = {}
        for key, v in self.state_dict().items():
            if not key.startswith("lora"):
                continue
            v = v.detach()
            if dtype is not None:
                v = v.to("cpu").to(dtype)
            sd[key] = v.contiguous()
        if os.path.splitext(file)[1] == ".safetensors":
            from safetensors.torch import save_file
            save_file(sd, file, metadata)
        else:
            torch.save(sd, file)

    def load_weights(self, file, strict: bool = True):
        """Inverse of save_weights (the reference has none; SURVEY §8f rank 3: needed for resume and for the
        notebook's inference cell).  Accepts the kohya / A1111 key set `<name>.alpha|lora_down.weight|lora_up.weight`
        from a .safetensors or torch file; values are copied IN PLACE so the flat operand buffer (and any CUDA graph
        that captured its address) stays valid."""
        if os.path.splitext(file)[1] == ".safetensors":
            from safetensors.torch import load_file
            sd = load_file(file)
        else:
            sd = torch.load(file, map_location="cpu")
        own = {k: v for k, v in self.state_dict().items() if k.startswith("lora")}
        missing = [k for k in own if k not in sd]
        unexpected = [k for k in sd if k not in own]
        if strict and (missing or unexpected):
            raise KeyError(f"load_weights: missing {missing[:3]}{'...' if len(missing) > 3 else ''}, "
                           f"unexpected {unexpected[:3]}{'...' if len(unexpected) > 3 else ''}")
        with torch.no_grad():
            for k, dst in own.items():
                if k in sd:
                    if tuple(sd[k].shape) != tuple(dst.shape):
                        raise ValueError(f"load_weights: {k} has shape {tuple(sd[k].shape)}, expected {tuple(dst.shape)}")
                    dst.copy_(sd[k].to(device=dst.device, dtype=dst.dtype))
        if self.flat is not None:
            self.flat.refresh_transposed()
        return missing, unexpected

    def __enter__(self):
        for lora in self.unet_loras:
            lora.multiplier = 1.0

    def __exit__(self, exc_type, exc_value, tb):
        for lora in self.unet_loras:
            lora.multiplier = 0

    # ---- flat storage ---------------------------------------------------------------------
    def _apply(self, fn, *a, **k):
        out = super()._apply(fn, *a, **k)
        self.flat = None
        p = next(iter(self.parameters()), None)
        if p is not None and p.is_cuda and self._engine:
            self.bind_flat()
        return out

    def bind_flat(self):
        """Re-home every adapter weight as a view of one flat buffer in engine-site layout."""
        eng = self._engine[0]
        p0 = next(self.parameters())
        eng._ensure_packed(p0.device)
        sites = [s for s in eng.lora_sites() if s.adapters() is not None]
        mine = {id(l) for l in self.unet_loras}
        offsets, total = [], 0
        for s in sites:
            ads = s.adapters()
            if any(id(a) not in mine for a in ads):
                raise RuntimeError("leco_b200: UNet is adapted by another network")
            kl = (sum(a.lora_down.weight.shape[0] for a in ads) + 15) // 16 * 16
            offsets.append((total, kl))
            total += kl * s.k_in + s.n_total * kl
        covered = sum(len(s.adapters()) for s in sites)
        if covered != len(self.unet_loras):
            raise NotImplementedError(
                f"leco_b200: {len(self.unet_loras) - covered} adapters sit on layers without a fused engine site")
        dev, dt = p0.device, p0.dtype
        if dt not in (torch.bfloat16, torch.float32):
            raise NotImplementedError(
                f"leco_b200: adapters in {dt} are not supported — use network.to(device, dtype=torch.bfloat16) "
                "(train.precision: bfloat16, every example config) or torch.float32 (fp32 master adapters); the "
                "reference's README calls float16 unstable")
        master = torch.zeros(total, device=dev, dtype=torch.float32) if dt == torch.float32 else None
        st = FlatState(torch.zeros(total, device=dev, dtype=torch.bfloat16),
                       torch.zeros(total + 8, device=dev, dtype=torch.float32),
                       torch.zeros(total, device=dev, dtype=torch.uint8), master=master)
        home = st.params if master is None else master      # where the Parameters live
        for s, (off, kl) in zip(sites, offsets):
            na, nb = kl * s.k_in, s.n_total * kl
            op_ad, op_bup = st.params[off:off + na].view(kl, s.k_in), st.params[off + na:off + na + nb].view(s.n_total, kl)
            ad, bup = home[off:off + na].view(kl, s.k_in), home[off + na:off + na + nb].view(s.n_total, kl)
            g_ad, g_bup = st.grads[off:off + na].view(kl, s.k_in), st.grads[off + na:off + na + nb].view(s.n_total, kl)
            m_ad, m_bup = st.mask[off:off + na].view(kl, s.k_in), st.mask[off + na:off + na + nb].view(s.n_total, kl)
            k0 = 0
            for a, n0 in zip(s.adapters(), s.n_offsets):
                wd, wu = a.lora_down.weight, a.lora_up.weight
                r, n = wd.shape[0], wu.shape[0]
                ad[k0:k0 + r].copy_(down_as_rows(wd.detach()))
                bup[n0:n0 + n, k0:k0 + r].copy_(wu.detach().reshape(n, r))
                wd.data = rows_as_down(ad[k0:k0 + r], wd.shape)   # 3x3 adapters: a permuted (OHWI-stored) view
                wu.data = bup[n0:n0 + n, k0:k0 + r].view(n, r, *wu.shape[2:]) if wu.dim() == 2 else \
                    bup[n0:n0 + n, k0:k0 + r].unsqueeze(-1).unsqueeze(-1)
                m_ad[k0:k0 + r] = 1
                m_bup[n0:n0 + n, k0:k0 + r] = 1
                k0 += r
            s.bind_native(op_ad, op_bup, g_ad, g_bup, home_ptr=ad.data_ptr())
        st.n_real = int(st.mask.sum().item())
        st.build_transposed(sites, offsets)
        self.flat = st
        return st

    def adapter_grads(self):
        """[d lora_down.weight, d lora_up.weight] per adapter in `prepare_optimizer_params` order, read out of the
        flat fp32 gradient buffer (views shaped like the parameters; valid until the optimizer zeroes the buffer)."""
        if self.flat is None:
            raise RuntimeError("adapter_grads needs the flat layout (network on CUDA, bound to an EngineUNet)")
        views = {}
        for s in self._engine[0].lora_sites():
            ads = s.adapters()
            if ads is None:
                continue
            k0 = 0
            for a, n0 in zip(ads, s.n_offsets):
                wd, wu = a.lora_down.weight, a.lora_up.weight
                r, n = wd.shape[0], wu.shape[0]
                views[id(a)] = (rows_as_down(s.g_ad[k0:k0 + r], wd.shape), s.g_bup[n0:n0 + n, k0:k0 + r].reshape(wu.shape))
                k0 += r
        out = []
        for lora in self.unet_loras:
            out += list(views[id(lora)])
        return out


class FlatState:
    """Flat adapter storage: bf16 params (the GEMM operands), fp32 grads, uint8 mask (1 = a real LoRA element, 0 =
    operand padding / off-block zero that must never be updated).  `grads` is a view of `grads_ext`, which carries one
    more fp32 slot (`loss_slot`) so that data-parallel training reduces gradient and loss in ONE all-reduce.
    `master` (fp32, same layout) exists for a float32 network: the Parameters are views of it and `params` is its
    bf16 copy."""

    def __init__(self, params, grads_ext, mask, master=None):
        n = params.numel()
        self.params, self.grads_ext, self.mask, self.master = params, grads_ext, mask, master
        self.grads = grads_ext[:n]
        self.loss_slot = grads_ext[n:n + 1]
        self.n_real = 0
        self.params_t = None      # flat buffer of the transposed operands (ad^T, bup^T per site)
        self._tiles = None
        self._n_tiles = 0

    def build_transposed(self, sites, offsets):
        """Static (ad^T [K,Kl], bup^T [Kl,N]) views per site + the 32x32 tile table of the one-launch refresh."""
        import struct
        self.params_t = torch.zeros_like(self.params)
        recs = []
        for s, (off, kl) in zip(sites, offsets):
            na, nb = kl * s.k_in, s.n_total * kl
            s.static_t = (self.params_t[off:off + na].view(s.k_in, kl), self.params_t[off + na:off + na + nb].view(kl, s.n_total))
            s.flat_state = self
            for o, rows, cols in ((off, kl, s.k_in), (off + na, s.n_total, kl)):
                for r0 in range(0, rows, 32):
                    for c0 in range(0, cols, 32):
                        recs.append(struct.pack("<qqiiii", o, o, rows, cols, r0, c0))
        self._n_tiles = len(recs)
        self._tiles = torch.frombuffer(bytearray(b"".join(recs)), dtype=torch.uint8).to(self.params.device)
        self.refresh_transposed()

    @property
    def home(self):
        """The buffer the Parameters are views of (what an optimizer updates, what save_weights reads)."""
        return self.params if self.master is None else self.master

    def refresh_operands(self):
        """float32 network: re-derive the bf16 operand buffer from the fp32 master (one launch).  A no-op for a bf16
        network, whose Parameters ARE the operands."""
        if self.master is not None:
            from . import ops
            ops.cast_f32_to_bf16(self.master, out=self.params)

    def refresh_transposed(self):
        """Bring every site's operands and (ad^T, bup^T) up to date with the parameters: one kernel launch (two for a
        float32 network).  Called before every pass that differentiates (LecoTrainer.iteration, EngineUNet.forward
        under autograd), so in-place parameter edits by any optimizer / loader are picked up."""
        self.refresh_operands()
        if self.params_t is not None:
            from . import ops
            ops.transpose_tiles(self.params, self.params_t, self._tiles, self._n_tiles)


OPTIMIZER_MODES = {"adamw": 0, "adam": 1, "lion": 2}


class FlatOptimizer:
    """The reference's optimizers (train_util.get_optimizer, train_util.py:333-370) that have a closed elementwise
    form — torch.optim.AdamW / torch.optim.Adam / lion_pytorch.Lion — as ONE fused kernel over the flat LoRA buffer.
    Optimizer state lives in the parameter dtype like the reference's (`network.to(dtype)` precedes the optimizer,
    train_lora.py:78-89) and the kernel then rounds where torch's foreach implementation rounds; `state_fp32=True`
    keeps fp32 moments instead (no intermediate rounding); a float32 network (fp32 master parameters) always has fp32
    moments and is stepped as torch steps fp32 tensors.  `param_groups[0]["lr"]` is read at every step, so any
    torch.optim.lr_scheduler-style driver works (train_lora.py:281)."""

    def __init__(self, flat: FlatState, name: str = "adamw", lr=1e-3, betas=None, eps=1e-8, weight_decay=None,
                 state_fp32: bool = False, **unknown):
        name = name.lower()
        if name not in OPTIMIZER_MODES:
            raise ValueError("Optimizer must be adam, adamw, lion or Prodigy")
        if unknown:
            raise TypeError(f"{name}: unexpected optimizer arguments {sorted(unknown)}")
        self.mode = OPTIMIZER_MODES[name]
        if betas is None:
            betas = (0.9, 0.99) if name == "lion" else (0.9, 0.999)
        if weight_decay is None:
            weight_decay = 1e-2 if name == "adamw" else 0.0
        self.flat = flat
        sd = torch.float32 if (state_fp32 or flat.master is not None) else flat.params.dtype
        self.exp_avg = torch.zeros_like(flat.params, dtype=sd)
        self.exp_avg_sq = torch.zeros_like(flat.params, dtype=sd) if self.mode != 2 else self.exp_avg
        self.step_count = 0
        self.defaults = dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay)
        self.param_groups = [dict(self.defaults, params=[flat.home])]
        self.hyper = torch.zeros(16, device=flat.params.device, dtype=torch.float32)

    @property
    def lr(self):
        return self.param_groups[0]["lr"]

    @lr.setter
    def lr(self, v):
        self.param_groups[0]["lr"] = v

    def step(self, grad_scale: float = 1.0):
        from . import ops
        self.step_count += 1
        g = self.param_groups[0]
        lr, (b1, b2), eps, wd = float(g["lr"]), g["betas"], g["eps"], g["weight_decay"]
        t = self.step_count
        # double-precision host scalars exactly as torch forms them (adam.py: bias corrections, step size)
        bc1, bc2 = 1 - b1 ** t, 1 - b2 ** t
        # pageable source: the copy is staged before this call returns, so the host values can change
        self.hyper.copy_(torch.tensor([lr, b1, b2, eps, wd, float(t), grad_scale, float(self.mode),
                                       lr / bc1, bc2 ** 0.5, 1 - lr * wd, 1 - b1, 1 - b2, 0, 0, 0], dtype=torch.float32))
        if self.flat.master is not None:     # float32 network: fp32 master + moments, bf16 operands rewritten in-pass
            ops.optim_flat_master(self.flat.master, self.flat.params, self.flat.grads, self.exp_avg, self.exp_avg_sq,
                                  self.flat.mask, self.hyper, zero_grad=True)
        else:
            ops.optim_flat(self.flat.params, self.flat.grads, self.exp_avg, self.exp_avg_sq, self.flat.mask, self.hyper,
                           zero_grad=True)

    def zero_grad(self):
        self.flat.grads.zero_()


class FlatAdamW(FlatOptimizer):
    """torch.optim.AdamW (train_lora.py:89, train_util.py:357-360): the example configs' optimizer."""

    def __init__(self, flat: FlatState, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2,
                 state_fp32: bool = False):
        super().__init__(flat, "adamw", lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, state_fp32=state_fp32)
