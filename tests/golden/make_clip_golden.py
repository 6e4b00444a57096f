"""Golden vectors of the text-encoder prologue, made by the reference's own dependency (transformers) in the build
container: token ids of CLIPTokenizer on the synthetic vocabulary of tests/clip_fixtures.py, and outputs of
random-initialised CLIPTextModel / CLIPTextModelWithProjection (fp32, weights rounded to bf16 first so the GPU engine
sees the very same numbers).  Writes tests/golden/clip_tokens.json and tests/golden/clip_tiny.pt.

    python tests/golden/make_clip_golden.py
"""
from __future__ import annotations

import json
import os
import sys
import tempfile

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from tests.clip_fixtures import PROMPTS, hf_config, token_ids_for, write_tokenizer_dir  # noqa: E402
from leco_b200.text_encoder import TEXT_SPECS  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    import transformers
    from transformers import CLIPTextConfig, CLIPTextModel, CLIPTextModelWithProjection, CLIPTokenizer
    out = {"transformers": transformers.__version__, "prompts": PROMPTS, "ids": {}}
    for pad in ("<|endoftext|>", "!"):
        d = write_tokenizer_dir(tempfile.mkdtemp(), pad_token=pad)
        tok = CLIPTokenizer(os.path.join(d, "vocab.json"), os.path.join(d, "merges.txt"), pad_token=pad)
        ids = tok(PROMPTS, padding="max_length", max_length=77, truncation=True, return_tensors="pt").input_ids
        out["ids"][pad] = ids.tolist()
    with open(os.path.join(HERE, "clip_tokens.json"), "w", encoding="utf-8") as f:
        json.dump(out, f)

    blob = {"transformers": transformers.__version__}
    for name in ("tiny_clip", "tiny_clip_proj"):
        spec = TEXT_SPECS[name]
        torch.manual_seed(7)
        cfg = CLIPTextConfig(**hf_config(spec))
        model = (CLIPTextModelWithProjection if spec.projection_dim else CLIPTextModel)(cfg).eval()
        with torch.no_grad():
            for n, p in model.named_parameters():            # livelier than the default init, and bf16-exact
                if "layer_norm" in n and n.endswith("weight"):
                    p.copy_(1.0 + 0.1 * torch.randn_like(p))
                elif p.dim() == 1:
                    p.copy_(0.05 * torch.randn_like(p))
                elif "embedding" in n:
                    p.copy_(0.1 * torch.randn_like(p))
                else:
                    p.copy_(torch.randn_like(p) * (1.5 * p.shape[1] ** -0.5))
                p.copy_(p.to(torch.bfloat16).float())
        ids = token_ids_for(spec, batch=3, seed=11)
        with torch.no_grad():
            o = model(ids, output_hidden_states=True)
        sd = {k: v.to(torch.bfloat16) for k, v in model.state_dict().items() if v.is_floating_point()}
        blob[name] = {"state_dict": sd, "ids": ids, "first": o[0].clone(), "last_hidden_state": o.last_hidden_state.clone(),
                      "penultimate": o.hidden_states[-2].clone(),
                      "text_embeds": o.text_embeds.clone() if spec.projection_dim else None}
    torch.save(blob, os.path.join(HERE, "clip_tiny.pt"))
    print("wrote", os.path.join(HERE, "clip_tokens.json"), os.path.join(HERE, "clip_tiny.pt"),
          os.path.getsize(os.path.join(HERE, "clip_tiny.pt")), "bytes")


if __name__ == "__main__":
    main()
