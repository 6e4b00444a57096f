"""ORACLE (test infrastructure, CPU, fp32): restatement of the CLIP text encoder the reference calls.

Only tests/, `__graft_entry__.smoke()` and bench.py's cpu_baseline leg may import this; the product never does.

What the reference runs (`/root/reference/train_util.py:73-74` text_encode, `:88-103` text_encode_xl) is transformers'
`CLIPTextModel` / `CLIPTextModelWithProjection` (third-party dependency, pinned `transformers==4.33.1` in
`/root/reference/requirements.txt`; not vendored).  Published algorithm, `modeling_clip.py`:

    CLIPTextEmbeddings      x = token_embedding[ids] + position_embedding[0..S)
    CLIPEncoderLayer  x L   x = x + out_proj(softmax(causal(q k^T d^-1/2)) v),  q,k,v = proj(LayerNorm1(x))
                            x = x + fc2(act(fc1(LayerNorm2(x)))),  act = quick_gelu x*sigmoid(1.702 x) | erf GELU
    final_layer_norm        last = LayerNorm(x)
    pooled                  last[b, eos position]  (eos_token_id == 2: argmax of the ids, else first id == eos_token_id)
    text_projection         text_embeds = pooled @ W^T            (WithProjection only)
    hidden_states           (embeddings, after layer 1, ..., after layer L)  -- text_encode_xl reads [-2]

PINNED: `tests/test_text_prologue_cpu.py::test_clip_oracle_matches_transformers` runs this against the transformers
package of this image (5.5, random-initialised CLIPTextModel / CLIPTextModelWithProjection of the same config, fp32,
max |diff| < 2e-5) for both activations and both pooling rules; `tests/golden/clip_tiny.pt` holds outputs made the
same way by `tests/golden/make_clip_golden.py`.  No real checkpoint exists offline: weights are always synthetic.
"""
from __future__ import annotations

from typing import Dict, Tuple

import torch
import torch.nn.functional as F


def clip_text_forward(sd: Dict[str, torch.Tensor], input_ids: torch.Tensor, *, heads: int, act: str = "quick_gelu",
                      eps: float = 1e-5, eos_token_id: int = 2, dtype=torch.float32
                      ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, tuple]:
    """-> (last_hidden_state [B,S,D], pooled [B,D], text_embeds [B,P] | None, hidden_states tuple of L+1 [B,S,D])."""
    w = {k: v.to(dtype) for k, v in sd.items() if v.is_floating_point()}
    ids = input_ids.long()
    B, S = ids.shape
    pre = "text_model."
    x = w[pre + "embeddings.token_embedding.weight"][ids] + w[pre + "embeddings.position_embedding.weight"][:S][None]
    D = x.shape[-1]
    d = D // heads
    mask = torch.full((S, S), float("-inf"), dtype=dtype, device=x.device).triu(1)
    hidden = [x]
    n_layers = 1 + max(int(k.split(".")[3]) for k in w if k.startswith(pre + "encoder.layers."))
    for i in range(n_layers):
        p = f"{pre}encoder.layers.{i}."

        def lin(t, name):
            return F.linear(t, w[p + name + ".weight"], w[p + name + ".bias"])
        y = F.layer_norm(x, (D,), w[p + "layer_norm1.weight"], w[p + "layer_norm1.bias"], eps)
        q, k, v = (lin(y, f"self_attn.{n}_proj").view(B, S, heads, d).transpose(1, 2) for n in "qkv")
        att = torch.softmax(q @ k.transpose(-1, -2) * d ** -0.5 + mask, dim=-1) @ v
        x = x + lin(att.transpose(1, 2).reshape(B, S, D), "self_attn.out_proj")
        y = lin(F.layer_norm(x, (D,), w[p + "layer_norm2.weight"], w[p + "layer_norm2.bias"], eps), "mlp.fc1")
        y = y * torch.sigmoid(1.702 * y) if act == "quick_gelu" else F.gelu(y)
        x = x + lin(y, "mlp.fc2")
        hidden.append(x)
    last = F.layer_norm(x, (D,), w[pre + "final_layer_norm.weight"], w[pre + "final_layer_norm.bias"], eps)
    pos = ids.argmax(-1) if eos_token_id == 2 else (ids == eos_token_id).int().argmax(-1)
    pooled = last[torch.arange(B), pos]
    proj = w.get("text_projection.weight")
    return last, pooled, (pooled @ proj.t() if proj is not None else None), tuple(hidden)
