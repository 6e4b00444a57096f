"""Shared helpers of the text-prologue tests: a synthetic CLIP vocabulary (no real one exists offline), the prompt
list, and transformers <-> leco_b200 config translation.  Test infrastructure only."""
from __future__ import annotations

import json
import os

from leco_b200.tokenizer import BOS, EOS, bytes_to_unicode

MERGES = [("t", "h"), ("th", "e</w>"), ("i", "n"), ("a", "n"), ("an", "d</w>"), ("c", "a"), ("ca", "t</w>"), ("d", "o"),
          ("do", "g</w>"), ("p", "h"), ("ph", "o"), ("pho", "t"), ("phot", "o</w>"), ("in", "g</w>"), ("p", "a"),
          ("pa", "in"), ("pain", "t"), ("paint", "ing</w>"), ("'", "s</w>"), ("a", "n</w>"), ("v", "an</w>"), ("g", "o"),
          ("go", "g"), ("gog", "h</w>"), ("!", "!</w>"), ("4", "k</w>"), ("Ã", "©"), ("o", "f</w>"), ("s", "t"),
          ("st", "y"), ("sty", "l"), ("styl", "e</w>")]

PROMPTS = ["a photo of the cat and dog", "Van Gogh's   painting, 4k!!", "", "  ", "café au lait — déjà vu",
           "it's 2023: I'm here; they'll've", "日本語 text", "x " * 100, "UPPER lower MiXeD", "tab\there\nnewline",
           "emoji 🙂 ok", "a.b,c!d?e", "<|endoftext|> inside", "123 4567", "the the the the", "café nfd",
           "IT'S Don'T", "van gogh style", "painting of a dog, photo of a cat"]


def synthetic_vocab():
    alphabet = list(bytes_to_unicode().values())
    vocab = alphabet + [c + "</w>" for c in alphabet] + [a + b for a, b in MERGES] + [BOS, EOS]
    return {t: i for i, t in enumerate(vocab)}          # 512 + len(MERGES) + 2 entries


def write_tokenizer_dir(d: str, pad_token: str = EOS) -> str:
    os.makedirs(d, exist_ok=True)
    with open(os.path.join(d, "vocab.json"), "w", encoding="utf-8") as f:
        json.dump(synthetic_vocab(), f)
    with open(os.path.join(d, "merges.txt"), "w", encoding="utf-8") as f:
        f.write("#version: 0.2\n" + "\n".join(f"{a} {b}" for a, b in MERGES) + "\n")
    with open(os.path.join(d, "tokenizer_config.json"), "w", encoding="utf-8") as f:
        json.dump({"model_max_length": 77, "pad_token": pad_token, "unk_token": EOS}, f)
    return d


def hf_config(spec):
    """leco_b200 ClipTextSpec -> transformers CLIPTextConfig kwargs."""
    kw = dict(vocab_size=spec.vocab_size, hidden_size=spec.hidden_size, intermediate_size=spec.intermediate_size,
              num_hidden_layers=spec.num_hidden_layers, num_attention_heads=spec.num_attention_heads,
              max_position_embeddings=spec.max_position_embeddings, hidden_act=spec.hidden_act,
              layer_norm_eps=spec.layer_norm_eps, eos_token_id=spec.eos_token_id, bos_token_id=spec.vocab_size - 2,
              pad_token_id=1)
    if spec.projection_dim:
        kw["projection_dim"] = spec.projection_dim
    return kw


def token_ids_for(spec, batch: int, seed: int = 0):
    """[batch, 77] ids shaped like real prompts: bos, a few words, eos, padding (eos id is the largest id)."""
    import torch
    g = torch.Generator().manual_seed(seed)
    eos = spec.vocab_size - 1 if spec.eos_token_id == 2 else spec.eos_token_id
    bos = eos - 1
    ids = torch.full((batch, 77), eos, dtype=torch.long)
    for b in range(batch):
        n = int(torch.randint(0, 70, (1,), generator=g))
        ids[b, 0] = bos
        ids[b, 1:1 + n] = torch.randint(0, bos, (n,), generator=g)
    return ids


def unet_config_json(spec) -> dict:
    """leco_b200 UNetSpec -> the `unet/config.json` a diffusers checkpoint of that topology carries."""
    levels = len(spec.block_out_channels)
    cfg = {"_class_name": "UNet2DConditionModel", "block_out_channels": list(spec.block_out_channels),
           "down_block_types": ["CrossAttnDownBlock2D" if a else "DownBlock2D" for a in spec.attn_levels],
           "up_block_types": ["CrossAttnUpBlock2D" if a else "UpBlock2D" for a in reversed(spec.attn_levels)],
           "layers_per_block": spec.layers_per_block, "cross_attention_dim": spec.cross_attention_dim,
           "attention_head_dim": list(spec.num_heads), "use_linear_projection": spec.use_linear_projection,
           "norm_num_groups": spec.norm_groups, "in_channels": spec.in_channels, "out_channels": spec.out_channels,
           "transformer_layers_per_block": list(spec.transformer_depth)}
    assert len(spec.num_heads) == levels
    if spec.text_time:
        cfg.update(addition_embed_type="text_time", addition_time_embed_dim=spec.add_time_dim,
                   projection_class_embeddings_input_dim=spec.add_proj_in)
    return cfg


def write_checkpoint_dir(root: str, arch: str, seed: int = 0, bin_format: bool = False) -> str:
    """A diffusers-layout checkpoint directory with synthetic weights: unet/ tokenizer/ text_encoder/ (and the _2 pair
    for text_time architectures), the text widths chosen to add up to the UNet's cross_attention_dim."""
    import torch
    from safetensors.torch import save_file
    from leco_b200.synthetic import build_engine
    from leco_b200.text_encoder import ClipTextEncoder, ClipTextSpec
    from leco_b200.unet import SPECS
    spec = SPECS[arch]
    os.makedirs(os.path.join(root, "unet"), exist_ok=True)
    with open(os.path.join(root, "unet", "config.json"), "w") as f:
        json.dump(unet_config_json(spec), f)
    unet = build_engine(arch, "cpu", seed=seed)
    sd = {k: v.detach().clone().contiguous() for k, v in unet.state_dict().items()}
    if bin_format:
        torch.save(sd, os.path.join(root, "unet", "diffusion_pytorch_model.bin"))
    else:
        save_file(sd, os.path.join(root, "unet", "diffusion_pytorch_model.safetensors"))
    vocab = len(synthetic_vocab())
    if spec.text_time:
        w2 = spec.add_text_dim
        encs = [("", ClipTextSpec("te1", vocab_size=vocab, hidden_size=spec.cross_attention_dim - w2, intermediate_size=128,
                                  num_hidden_layers=3, num_attention_heads=2, eos_token_id=2), EOS, "CLIPTextModel"),
                ("_2", ClipTextSpec("te2", vocab_size=vocab, hidden_size=w2, intermediate_size=128, num_hidden_layers=3,
                                    num_attention_heads=2, hidden_act="gelu", projection_dim=w2, eos_token_id=2), "!",
                 "CLIPTextModelWithProjection")]
    else:
        encs = [("", ClipTextSpec("te", vocab_size=vocab, hidden_size=spec.cross_attention_dim, intermediate_size=256,
                                  num_hidden_layers=2, num_attention_heads=2, hidden_act="gelu" if spec.use_linear_projection
                                  else "quick_gelu", eos_token_id=2), "!" if spec.use_linear_projection else EOS,
                 "CLIPTextModel")]
    for suffix, ts, pad, cls_name in encs:
        write_tokenizer_dir(os.path.join(root, "tokenizer" + suffix), pad_token=pad)
        d = os.path.join(root, "text_encoder" + suffix)
        os.makedirs(d, exist_ok=True)
        cfg = hf_config(ts)
        cfg["architectures"] = [cls_name]
        cfg.setdefault("projection_dim", 512)            # CLIPTextModel configs carry one too; it must be ignored
        with open(os.path.join(d, "config.json"), "w") as f:
            json.dump(cfg, f)
        torch.manual_seed(seed + len(suffix))
        enc = ClipTextEncoder(ts)
        with torch.no_grad():
            for n, p in enc.named_parameters():
                if "layer_norm" in n and n.endswith("weight"):
                    p.copy_(1.0 + 0.1 * torch.randn_like(p))
                elif p.dim() == 1:
                    p.copy_(0.05 * torch.randn_like(p))
                else:
                    p.copy_(torch.randn_like(p) * (0.1 if "embedding" in n else 1.5 * p.shape[1] ** -0.5))
                p.copy_(p.to(torch.bfloat16).float())
        esd = {k: v.detach().clone().contiguous() for k, v in enc.state_dict().items()}
        esd["text_model.embeddings.position_ids"] = torch.arange(77)[None]          # as old checkpoints carry it
        save_file(esd, os.path.join(d, "model.safetensors"))
    return root
