"""CPU suite of the text-encoder prologue: the tokenizer against transformers' CLIPTokenizer and
the committed golden ids, the oracle restatement against transformers' CLIPTextModel, and the engine encoder's wiring
(parameter names, q/k/v stacking, causal mask, pooling rule, hidden_states[-2]) through the plain-torch test double."""
import json
import os
import tempfile

import pytest
import torch

from leco_b200.text_encoder import TEXT_SPECS, ClipTextEncoder, ClipTextSpec, build_text_encoder
from leco_b200.tokenizer import ClipTokenizer
from oracle import clip_ref
from tests import torch_backend
from tests.clip_fixtures import PROMPTS, hf_config, token_ids_for, write_tokenizer_dir

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("pad", ["<|endoftext|>", "!"])
def test_tokenizer_matches_golden_ids(pad):
    with open(os.path.join(GOLDEN, "clip_tokens.json"), encoding="utf-8") as f:
        gold = json.load(f)
    tok = ClipTokenizer.from_pretrained(write_tokenizer_dir(tempfile.mkdtemp(), pad_token=pad))
    assert tok.pad_token == pad and tok.model_max_length == 77
    ids = tok(gold["prompts"], padding="max_length", max_length=tok.model_max_length, truncation=True,
              return_tensors="pt").input_ids                                   # the call of train_util.py:64-70
    assert ids.shape == (len(gold["prompts"]), 77) and ids.dtype == torch.long
    assert ids.tolist() == gold["ids"][pad]


def test_tokenizer_matches_transformers_live():
    transformers = pytest.importorskip("transformers")
    d = write_tokenizer_dir(tempfile.mkdtemp())
    ref = transformers.CLIPTokenizer(os.path.join(d, "vocab.json"), os.path.join(d, "merges.txt"))
    mine = ClipTokenizer.from_pretrained(d)
    extra = ["a" * 300, "many , , , commas ,,, and... dots", "'s 't 're 've 'm 'll 'd", "MiXed'S CASE'LL", "1a2b3c"]
    for p in PROMPTS + extra:
        want = ref([p], padding="max_length", max_length=77, truncation=True, return_tensors="pt").input_ids[0].tolist()
        assert mine.encode(p) == want, p


def test_tokenizer_edges():
    tok = ClipTokenizer.from_pretrained(write_tokenizer_dir(tempfile.mkdtemp()))
    empty = tok.encode("")
    assert empty[:2] == [tok.bos_token_id, tok.eos_token_id] and set(empty[2:]) == {tok.pad_token_id}
    long = tok.encode("the " * 200)
    assert len(long) == 77 and long[0] == tok.bos_token_id and long[-1] == tok.eos_token_id     # truncated, eos kept
    assert tok("one string").input_ids.shape == (1, 77)
    with pytest.raises(ValueError):
        ClipTokenizer({"<|startoftext|>": 0, "<|endoftext|>": 1}, [], pad_token="!")


def _hf_model(spec):
    transformers = pytest.importorskip("transformers")
    cfg = transformers.CLIPTextConfig(**hf_config(spec))
    torch.manual_seed(3)
    cls = transformers.CLIPTextModelWithProjection if spec.projection_dim else transformers.CLIPTextModel
    model = cls(cfg).eval()
    with torch.no_grad():
        for n, p in model.named_parameters():
            if p.dim() == 1 and "layer_norm" not in n:
                p.copy_(0.05 * torch.randn_like(p))
            elif p.dim() == 2:
                p.copy_(torch.randn_like(p) * (0.1 if "embedding" in n else 1.5 * p.shape[1] ** -0.5))
    return model


@pytest.mark.parametrize("name", ["tiny_clip", "tiny_clip_proj"])
def test_clip_oracle_matches_transformers(name):
    """Pins oracle/clip_ref.py against the transformers package itself (both activations, both pooling rules)."""
    spec = TEXT_SPECS[name]
    model = _hf_model(spec)
    ids = token_ids_for(spec, batch=4, seed=5)
    with torch.no_grad():
        o = model(ids, output_hidden_states=True)
    last, pooled, emb, hidden = clip_ref.clip_text_forward(model.state_dict(), ids, heads=spec.num_attention_heads,
                                                           act=spec.hidden_act, eps=spec.layer_norm_eps,
                                                           eos_token_id=spec.eos_token_id)
    assert (last - o.last_hidden_state).abs().max() < 2e-5
    assert len(hidden) == len(o.hidden_states) == spec.num_hidden_layers + 1
    assert (hidden[-2] - o.hidden_states[-2]).abs().max() < 2e-5
    if spec.projection_dim:
        assert (emb - o.text_embeds).abs().max() < 2e-5 and (o[0] - o.text_embeds).abs().max() == 0
    else:
        assert (pooled - o.pooler_output).abs().max() < 2e-5 and emb is None


@pytest.mark.parametrize("name", ["tiny_clip", "tiny_clip_proj"])
def test_clip_oracle_matches_golden(name):
    blob = torch.load(os.path.join(GOLDEN, "clip_tiny.pt"), weights_only=False)[name]
    spec = TEXT_SPECS[name]
    last, _, emb, hidden = clip_ref.clip_text_forward(blob["state_dict"], blob["ids"], heads=spec.num_attention_heads,
                                                      act=spec.hidden_act, eos_token_id=spec.eos_token_id)
    assert (last - blob["last_hidden_state"]).abs().max() < 2e-5
    assert (hidden[-2] - blob["penultimate"]).abs().max() < 2e-5
    if spec.projection_dim:
        assert (emb - blob["text_embeds"]).abs().max() < 2e-5


@pytest.mark.parametrize("name", ["tiny_clip", "tiny_clip_proj"])
def test_engine_text_encoder_wiring_matches_golden(name):
    """The product module, fp32, on the torch double: transformers' state dict loads key for key and the outputs the
    reference reads (`[0]`, `.hidden_states[-2]`) equal transformers' own."""
    blob = torch.load(os.path.join(GOLDEN, "clip_tiny.pt"), weights_only=False)[name]
    enc = ClipTextEncoder(TEXT_SPECS[name], backend=torch_backend).float()
    missing = enc.load_state_dict({k: v.float() for k, v in blob["state_dict"].items()}, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    enc.compute_dtype = torch.float32
    out = enc(blob["ids"], output_hidden_states=True)
    assert (out[0].float() - blob["first"]).abs().max() < 5e-5
    assert (out.last_hidden_state.float() - blob["last_hidden_state"]).abs().max() < 5e-5
    assert (out.hidden_states[-2].float() - blob["penultimate"]).abs().max() < 5e-5
    assert len(out.hidden_states) == TEXT_SPECS[name].num_hidden_layers + 1


def test_text_spec_from_checkpoint_config_and_position_ids_key():
    cfg = {"architectures": ["CLIPTextModelWithProjection"], "vocab_size": 600, "hidden_size": 128,
           "intermediate_size": 256, "num_hidden_layers": 2, "num_attention_heads": 4, "hidden_act": "gelu",
           "projection_dim": 64, "eos_token_id": 2}
    spec = ClipTextSpec.from_config(cfg)
    assert spec.projection_dim == 64 and spec.hidden_act == "gelu" and spec.max_position_embeddings == 77
    cfg["architectures"] = ["CLIPTextModel"]
    assert ClipTextSpec.from_config(cfg).projection_dim == 0        # CLIPTextModel configs carry projection_dim too
    enc = ClipTextEncoder(TEXT_SPECS["tiny_clip"], backend=torch_backend)
    sd = dict(enc.state_dict())
    sd["text_model.embeddings.position_ids"] = torch.arange(77)[None]      # saved by transformers < 4.31
    enc.load_state_dict(sd)
    assert TEXT_SPECS["clip_l"].hidden_size == 768 and TEXT_SPECS["openclip_h"].num_hidden_layers == 23
    assert TEXT_SPECS["openclip_bigg"].projection_dim == 1280


def test_pooled_row_rules():
    enc = build_text_encoder("tiny_clip", device="cpu", backend=torch_backend)     # eos id 599: first occurrence
    ids = torch.tensor([[598, 5, 599, 599, 599], [598, 599, 7, 7, 599]])
    assert enc.eos_rows(ids).tolist() == [2, 1]
    legacy = build_text_encoder("tiny_clip_proj", device="cpu", backend=torch_backend)   # eos id 2: argmax of ids
    assert legacy.eos_rows(torch.tensor([[598, 5, 599, 3, 3], [1, 2, 3, 4, 0]])).tolist() == [2, 3]


@pytest.mark.parametrize("arch,bin_format", [("tiny21", False), ("tiny15", True), ("tinyxl", False)])
def test_load_models_from_checkpoint_directory(arch, bin_format):
    """model_util.load_models[_xl] on a diffusers-layout directory: topology from unet/config.json, weights key for key,
    real tokenizer(s) + text encoder(s), and encode_prompts[_xl] equal to the oracle on the same tokens."""
    from leco_b200 import model_util
    from leco_b200.synthetic import build_engine
    from leco_b200.unet import SPECS
    from tests.clip_fixtures import write_checkpoint_dir
    d = write_checkpoint_dir(tempfile.mkdtemp(), arch, seed=3, bin_format=bin_format)
    xl = SPECS[arch].text_time
    if xl:
        toks, encs, unet, sched = model_util.load_models_xl(d, device="cpu")
    else:
        tok, enc, unet, sched = model_util.load_models(d, v2=arch == "tiny21", device="cpu")
        toks, encs = [tok], [enc]
    want = SPECS[arch]
    for field in ("block_out_channels", "attn_levels", "layers_per_block", "cross_attention_dim", "num_heads",
                  "transformer_depth", "use_linear_projection", "norm_groups", "text_time", "add_time_dim", "add_proj_in"):
        assert tuple(getattr(unet.spec, field)) == tuple(getattr(want, field)) if isinstance(
            getattr(want, field), (list, tuple)) else getattr(unet.spec, field) == getattr(want, field), field
    ref = build_engine(arch, "cpu", seed=3).state_dict()
    got = unet.state_dict()
    assert set(ref) == set(got) and all(torch.equal(ref[k], got[k]) for k in ref)
    assert [t.pad_token for t in toks] == (["<|endoftext|>", "!"] if xl else ["!" if arch == "tiny21" else "<|endoftext|>"])
    for e in encs:
        e._backend, e.compute_dtype = torch_backend, torch.float32
    prompts = ["van gogh style painting!"]
    if xl:
        out = model_util.encode_prompts_xl(toks, encs, prompts)
        text, pooled = out.text_embeds, out.pooled_embeds
        assert text.shape == (1, 77, want.cross_attention_dim) and pooled.shape == (1, want.add_text_dim)
    else:
        text = model_util.encode_prompts(toks[0], encs[0], prompts)
        assert text.shape == (1, 77, want.cross_attention_dim)
    parts = []
    for t, e in zip(toks, encs):
        ids = model_util.text_tokenize(t, prompts)
        last, _, emb, hidden = clip_ref.clip_text_forward(e.state_dict(), ids, heads=e.spec.num_attention_heads,
                                                          act=e.spec.hidden_act, eos_token_id=e.spec.eos_token_id)
        parts.append(hidden[-2] if xl else last)
    assert (torch.cat(parts, -1) - text.float()).abs().max() < 5e-5
    if xl:
        assert (emb - pooled.float()).abs().max() < 5e-5
