"""TEST FIXTURES: single-file checkpoints in the ORIGINAL Stable Diffusion (LDM) key layout with synthetic weights.

Written from the diffusers side, independently of leco_b200/ckpt_convert.py (which goes the other way): the UNet keys
are produced by enumerating the topology (block index arithmetic of openaimodel.UNetModel), the SD2.x / SDXL second
text tower in the open_clip layout (stacked in_proj, ln_1/ln_2, c_fc/c_proj, `x @ text_projection`).  The converter
must turn the file back into exactly the state dicts it was made from."""
import os
import re

import torch

_RES_INV = {"norm1": "in_layers.0", "conv1": "in_layers.2", "time_emb_proj": "emb_layers.1", "norm2": "out_layers.0",
            "conv2": "out_layers.3", "conv_shortcut": "skip_connection"}


def _res(tail: str) -> str:
    name, _, rest = tail.partition(".")
    return f"{_RES_INV[name]}.{rest}"


def unet_to_ldm(sd, spec):
    """diffusers UNet2DConditionModel keys -> `model.diffusion_model.*` (original layout)."""
    per = spec.layers_per_block + 1
    up_attn = list(reversed(spec.attn_levels))
    out = {}
    for k, v in sd.items():
        m = re.match(r"(down|up)_blocks\.(\d+)\.(resnets|attentions|downsamplers|upsamplers)\.(\d+)\.(.+)", k)
        if m:
            side, b, kind, l, tail = m.group(1), int(m.group(2)), m.group(3), int(m.group(4)), m.group(5)
            if side == "down":
                if kind == "resnets":
                    nk = f"input_blocks.{1 + b * per + l}.0.{_res(tail)}"
                elif kind == "attentions":
                    nk = f"input_blocks.{1 + b * per + l}.1.{tail}"
                else:
                    assert tail.startswith("conv.")
                    nk = f"input_blocks.{1 + b * per + spec.layers_per_block}.0.op.{tail[5:]}"
            else:
                if kind == "resnets":
                    nk = f"output_blocks.{b * per + l}.0.{_res(tail)}"
                elif kind == "attentions":
                    nk = f"output_blocks.{b * per + l}.1.{tail}"
                else:
                    nk = f"output_blocks.{b * per + spec.layers_per_block}.{2 if up_attn[b] else 1}.{tail}"
        elif k.startswith("mid_block.resnets."):
            n, tail = k[len("mid_block.resnets."):].split(".", 1)
            nk = f"middle_block.{0 if n == '0' else 2}.{_res(tail)}"
        elif k.startswith("mid_block.attentions.0."):
            nk = "middle_block.1." + k[len("mid_block.attentions.0."):]
        elif k.startswith("time_embedding.linear_"):
            n, tail = k[len("time_embedding.linear_"):].split(".", 1)
            nk = f"time_embed.{0 if n == '1' else 2}.{tail}"
        elif k.startswith("add_embedding.linear_"):
            n, tail = k[len("add_embedding.linear_"):].split(".", 1)
            nk = f"label_emb.0.{0 if n == '1' else 2}.{tail}"
        elif k.startswith("conv_in."):
            nk = "input_blocks.0.0." + k[len("conv_in."):]
        elif k.startswith("conv_norm_out."):
            nk = "out.0." + k[len("conv_norm_out."):]
        elif k.startswith("conv_out."):
            nk = "out.2." + k[len("conv_out."):]
        else:
            raise KeyError(k)
        assert nk not in out, nk
        out["model.diffusion_model." + nk] = v
    return out


def clip_to_open_clip(sd, prefix: str, extra_block: bool, generator=None):
    """transformers CLIPTextModel[WithProjection] keys -> an open_clip text tower under `prefix`; `extra_block` appends
    one more transformer block (SD2.x files carry 24, diffusers' text encoder keeps 23)."""
    out = {prefix + "positional_embedding": sd["text_model.embeddings.position_embedding.weight"],
           prefix + "token_embedding.weight": sd["text_model.embeddings.token_embedding.weight"],
           prefix + "ln_final.weight": sd["text_model.final_layer_norm.weight"],
           prefix + "ln_final.bias": sd["text_model.final_layer_norm.bias"],
           prefix + "logit_scale": torch.tensor(4.6052)}
    d = sd["text_model.embeddings.token_embedding.weight"].shape[1]
    out[prefix + "text_projection"] = (sd["text_projection.weight"].t().contiguous() if "text_projection.weight" in sd
                                       else torch.randn(d, d, generator=generator))
    n_layers = 1 + max(int(m.group(1)) for m in (re.match(r"text_model\.encoder\.layers\.(\d+)\.", k) for k in sd) if m)
    for n in range(n_layers + int(extra_block)):
        src = f"text_model.encoder.layers.{min(n, n_layers - 1)}."
        dst = f"{prefix}transformer.resblocks.{n}."

        def g(name):
            t = sd[src + name]
            return t if n < n_layers else torch.randn(t.shape, generator=generator)      # the block that gets dropped
        for kind in ("weight", "bias"):
            out[dst + "attn.in_proj_" + kind] = torch.cat([g(f"self_attn.{p}_proj.{kind}") for p in "qkv"], 0)
            out[dst + "attn.out_proj." + kind] = g("self_attn.out_proj." + kind)
            out[dst + "ln_1." + kind] = g("layer_norm1." + kind)
            out[dst + "ln_2." + kind] = g("layer_norm2." + kind)
            out[dst + "mlp.c_fc." + kind] = g("mlp.fc1." + kind)
            out[dst + "mlp.c_proj." + kind] = g("mlp.fc2." + kind)
    return out


def _encoder(spec, seed):
    from leco_b200.text_encoder import ClipTextEncoder
    torch.manual_seed(seed)
    enc = ClipTextEncoder(spec)
    with torch.no_grad():
        for n, p in enc.named_parameters():
            if "layer_norm" in n and n.endswith("weight"):
                p.copy_(1.0 + 0.1 * torch.randn_like(p))
            elif p.dim() == 1:
                p.copy_(0.05 * torch.randn_like(p))
            else:
                p.copy_(torch.randn_like(p) * (0.1 if "embedding" in n else 1.5 * p.shape[1] ** -0.5))
    return {k: v.detach().clone() for k, v in enc.state_dict().items()}


def write_single_file(root: str, arch: str, seed: int = 0, ext: str = ".safetensors"):
    """-> (path, unet state dict, [text-encoder state dicts]) — an LDM-layout file of the architecture's family
    (tiny15 -> SD1.x, tiny21 -> SD2.x, tinyxl -> SDXL) with `tokenizer/` [`tokenizer_2/`] beside it."""
    from leco_b200.synthetic import build_engine
    from leco_b200.text_encoder import ClipTextSpec
    from leco_b200.unet import SPECS
    from tests.clip_fixtures import EOS, synthetic_vocab, write_tokenizer_dir
    spec = SPECS[arch]
    os.makedirs(root, exist_ok=True)
    unet_sd = {k: v.detach().clone() for k, v in build_engine(arch, "cpu", seed=seed).state_dict().items()}
    vocab = len(synthetic_vocab())
    g = torch.Generator().manual_seed(seed + 100)

    def tspec(name, hidden, layers, act, proj=0):
        return ClipTextSpec(name, vocab_size=vocab, hidden_size=hidden, intermediate_size=2 * hidden, num_hidden_layers=layers,
                            num_attention_heads=max(1, hidden // 64), hidden_act=act, projection_dim=proj, eos_token_id=2)
    sd = unet_to_ldm(unet_sd, spec)
    if spec.text_time:
        w2 = spec.add_text_dim
        te = [_encoder(tspec("l", spec.cross_attention_dim - w2, 3, "quick_gelu"), seed + 1),
              _encoder(tspec("g", w2, 3, "gelu", proj=w2), seed + 2)]
        sd.update({"conditioner.embedders.0.transformer." + k: v for k, v in te[0].items()})
        sd["conditioner.embedders.0.transformer.text_model.embeddings.position_ids"] = torch.arange(77)[None]
        sd.update(clip_to_open_clip(te[1], "conditioner.embedders.1.model.", extra_block=False, generator=g))
        write_tokenizer_dir(os.path.join(root, "tokenizer"), pad_token=EOS)
        write_tokenizer_dir(os.path.join(root, "tokenizer_2"), pad_token="!")
    elif spec.use_linear_projection:      # SD2.x family
        te = [_encoder(tspec("h", spec.cross_attention_dim, 3, "gelu"), seed + 1)]
        sd.update(clip_to_open_clip(te[0], "cond_stage_model.model.", extra_block=True, generator=g))
        write_tokenizer_dir(os.path.join(root, "tokenizer"), pad_token="!")
    else:
        te = [_encoder(tspec("l", spec.cross_attention_dim, 2, "quick_gelu"), seed + 1)]
        sd.update({"cond_stage_model.transformer." + k: v for k, v in te[0].items()})
        write_tokenizer_dir(os.path.join(root, "tokenizer"), pad_token=EOS)
    # what every real file also carries and the loader must ignore: the VAE and the diffusion schedule buffers
    sd["first_stage_model.encoder.conv_in.weight"] = torch.randn(8, 3, 3, 3, generator=g)
    sd["alphas_cumprod"] = torch.linspace(0.99, 0.01, 10)
    sd = {k: v.contiguous() for k, v in sd.items()}
    path = os.path.join(root, "model" + ext)
    if ext == ".safetensors":
        from safetensors.torch import save_file
        save_file(sd, path)
    else:
        torch.save({"state_dict": sd, "global_step": 1}, path)
    return path, unet_sd, te
