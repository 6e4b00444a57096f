"""CPU: single-file checkpoints in the original Stable Diffusion key layout (the reference's `from_single_file` branch,
/root/reference/model_util.py:78-101, 111-118, 173-197, 207-216) load into the engine's parameter trees.

The conversion restates a third-party format (diffusers is absent: parity unpinned); what IS checked: a few key pairs of
the real SD1.5 / SDXL layouts written here as literals, and that a file written from the diffusers side by
tests/ldm_fixtures.py (independent index arithmetic) converts back to exactly the tensors it was made from — UNet,
text encoder(s) (open_clip towers incl. the dropped SD2.x block and the transposed projection), tokenizers."""
import tempfile

import pytest
import torch

from leco_b200 import ckpt_convert, model_util
from leco_b200.unet import SPECS
from oracle import clip_ref
from tests import torch_backend
from tests.ldm_fixtures import unet_to_ldm, write_single_file

# (original-layout key, diffusers key) pairs of published checkpoints
KNOWN_SD15 = [
    ("input_blocks.0.0.weight", "conv_in.weight"),
    ("time_embed.2.bias", "time_embedding.linear_2.bias"),
    ("input_blocks.1.1.transformer_blocks.0.attn1.to_q.weight", "down_blocks.0.attentions.0.transformer_blocks.0.attn1.to_q.weight"),
    ("input_blocks.2.0.emb_layers.1.weight", "down_blocks.0.resnets.1.time_emb_proj.weight"),
    ("input_blocks.3.0.op.weight", "down_blocks.0.downsamplers.0.conv.weight"),
    ("input_blocks.4.0.skip_connection.weight", "down_blocks.1.resnets.0.conv_shortcut.weight"),
    ("input_blocks.8.1.proj_out.bias", "down_blocks.2.attentions.1.proj_out.bias"),
    ("input_blocks.11.0.in_layers.2.weight", "down_blocks.3.resnets.1.conv1.weight"),
    ("middle_block.1.transformer_blocks.0.ff.net.0.proj.weight", "mid_block.attentions.0.transformer_blocks.0.ff.net.0.proj.weight"),
    ("middle_block.2.out_layers.3.weight", "mid_block.resnets.1.conv2.weight"),
    ("output_blocks.2.1.conv.weight", "up_blocks.0.upsamplers.0.conv.weight"),
    ("output_blocks.3.1.norm.weight", "up_blocks.1.attentions.0.norm.weight"),
    ("output_blocks.5.2.conv.bias", "up_blocks.1.upsamplers.0.conv.bias"),
    ("output_blocks.11.1.transformer_blocks.0.attn2.to_k.weight", "up_blocks.3.attentions.2.transformer_blocks.0.attn2.to_k.weight"),
    ("output_blocks.11.0.out_layers.0.weight", "up_blocks.3.resnets.2.norm2.weight"),
    ("out.0.weight", "conv_norm_out.weight"),
    ("out.2.bias", "conv_out.bias"),
]
KNOWN_SDXL = [
    ("label_emb.0.0.weight", "add_embedding.linear_1.weight"),
    ("label_emb.0.2.bias", "add_embedding.linear_2.bias"),
    ("input_blocks.2.0.in_layers.0.weight", "down_blocks.0.resnets.1.norm1.weight"),           # level 0: no attention
    ("input_blocks.4.1.transformer_blocks.1.attn2.to_v.weight", "down_blocks.1.attentions.0.transformer_blocks.1.attn2.to_v.weight"),
    ("input_blocks.8.1.transformer_blocks.2.norm3.weight", "down_blocks.2.attentions.1.transformer_blocks.2.norm3.weight"),   # twin: depth 3 (SDXL: 10)
    ("output_blocks.2.2.conv.weight", "up_blocks.0.upsamplers.0.conv.weight"),
    ("output_blocks.5.2.conv.weight", "up_blocks.1.upsamplers.0.conv.weight"),
    ("output_blocks.8.0.skip_connection.bias", "up_blocks.2.resnets.2.conv_shortcut.bias"),
]


@pytest.mark.parametrize("arch,known", [("tiny15", KNOWN_SD15), ("tinyxl", KNOWN_SDXL)])
def test_known_key_pairs_of_the_published_layouts(arch, known):
    from leco_b200.synthetic import build_engine
    sd = build_engine(arch, "cpu", seed=1).state_dict()
    ldm = unet_to_ldm(sd, SPECS[arch])
    back = ckpt_convert.convert_ldm_unet(ldm, SPECS[arch].layers_per_block)
    for a, b in known:
        assert ckpt_convert.UNET_PREFIX + a in ldm, a
        assert back[b] is ldm[ckpt_convert.UNET_PREFIX + a], (a, b)       # that very tensor, under the diffusers name
    if arch == "tiny15":   # SD1.x UNets have 12 input blocks, 12 output blocks, 686 tensors
        assert max(int(k.split(".")[3]) for k in ldm if ".input_blocks." in k) == 11
        assert max(int(k.split(".")[3]) for k in ldm if ".output_blocks." in k) == 11 and len(ldm) == 686
    with pytest.raises(KeyError):
        ckpt_convert.convert_ldm_unet({ckpt_convert.UNET_PREFIX + "bogus.0.weight": torch.zeros(1)})
    with pytest.raises(ValueError):
        ckpt_convert.detect_layout({"something.else": torch.zeros(1)})


@pytest.mark.parametrize("arch,ext", [("tiny15", ".safetensors"), ("tiny21", ".ckpt"), ("tinyxl", ".safetensors")])
def test_load_models_from_a_single_file_checkpoint(arch, ext):
    """model_util.load_models[_xl]("<file>.ckpt|.safetensors") = the reference's single-file branch."""
    root = tempfile.mkdtemp()
    path, unet_sd, te_sds = write_single_file(root, arch, seed=4, ext=ext)
    assert ckpt_convert.detect_layout(ckpt_convert.read_single_file(path)) == \
        {"tiny15": "ldm_sd1", "tiny21": "ldm_sd2", "tinyxl": "ldm_sdxl"}[arch]
    xl = SPECS[arch].text_time
    if xl:
        toks, encs, unet, _ = model_util.load_models_xl(path, device="cpu", arch=arch)
    else:
        tok, enc, unet, sched = model_util.load_models(path, v2=arch == "tiny21", v_pred=arch == "tiny21", device="cpu", arch=arch)
        toks, encs = [tok], [enc]
        assert sched.config.prediction_type == ("v_prediction" if arch == "tiny21" else "epsilon")
    got = unet.state_dict()
    assert set(got) == set(unet_sd) and all(torch.equal(got[k], unet_sd[k]) for k in got)
    assert len(encs) == len(te_sds) == (2 if xl else 1)
    for e, want in zip(encs, te_sds):
        have = e.state_dict()
        assert set(have) == set(want), set(have) ^ set(want)
        assert all(torch.equal(have[k], want[k]) for k in want)
        assert e.spec.num_attention_heads == max(1, e.spec.hidden_size // 64)
    assert [e.spec.hidden_act for e in encs] == (["quick_gelu", "gelu"] if xl else ["gelu" if arch == "tiny21" else "quick_gelu"])
    assert [t.pad_token for t in toks] == (["<|endoftext|>", "!"] if xl else ["!" if arch == "tiny21" else "<|endoftext|>"])
    if arch == "tiny21":
        assert encs[0].spec.num_hidden_layers == 3          # the file held 4 blocks; the last one is dropped
    # the prompt-encoding prologue runs on what was loaded and equals the oracle on the same tokens
    for e in encs:
        e._backend, e.compute_dtype = torch_backend, torch.float32
    prompts = ["van gogh style painting!"]
    if xl:
        out = model_util.encode_prompts_xl(toks, encs, prompts)
        text, pooled = out.text_embeds, out.pooled_embeds
        assert pooled.shape == (1, SPECS[arch].add_text_dim)
    else:
        text = model_util.encode_prompts(toks[0], encs[0], prompts)
    assert text.shape == (1, 77, SPECS[arch].cross_attention_dim)
    parts = []
    for t, e in zip(toks, encs):
        ids = model_util.text_tokenize(t, prompts)
        last, _, emb, hidden = clip_ref.clip_text_forward(e.state_dict(), ids, heads=e.spec.num_attention_heads,
                                                          act=e.spec.hidden_act, eos_token_id=e.spec.eos_token_id)
        parts.append(hidden[-2] if xl else last)
    assert (torch.cat(parts, -1) - text.float()).abs().max() < 5e-5
    if xl:
        assert (emb - pooled.float()).abs().max() < 5e-5


def test_single_file_error_behaviour(tmp_path):
    import os
    import shutil
    path, _, _ = write_single_file(str(tmp_path / "a"), "tiny15", seed=2)
    shutil.rmtree(os.path.join(os.path.dirname(path), "tokenizer"))
    with pytest.raises(FileNotFoundError, match="tokenizer/vocab.json"):          # the file holds no vocabulary
        model_util.load_models(path, device="cpu", arch="tiny15")
    with pytest.raises(ValueError, match="load_models_xl"):                       # SD1.x file into the XL loop
        path2, _, _ = write_single_file(str(tmp_path / "b"), "tiny15", seed=2)
        model_util._single_file_text_side(path2, "cpu", 2)
    # an old-style .ckpt that pickles foreign objects is refused unless explicitly trusted
    import argparse
    bad = str(tmp_path / "bad.ckpt")
    torch.save({"state_dict": {}, "hparams": argparse.Namespace(lr=1e-4)}, bad)
    with pytest.raises(RuntimeError, match="LECO_TRUST_CKPT"):
        ckpt_convert.read_single_file(bad)
    # a bare diffusers-layout UNet file keeps working (UNet real, stand-in text side)
    from safetensors.torch import save_file
    from leco_b200.synthetic import build_engine
    f = str(tmp_path / "unet_only.safetensors")
    save_file({k: v.contiguous() for k, v in build_engine("tiny21", "cpu", seed=5).state_dict().items()}, f)
    tok, enc, unet, _ = model_util.load_models(f, v2=True, device="cpu", arch="tiny21")
    assert tok is None and isinstance(enc, model_util.SyntheticTextEncoder)
