"""Generates the committed golden fixtures by running the reference's OWN unmodified code.

Runs only in the build container (needs /root/reference).  What it does:
  1. imports the reference's lora.py / train_util.py / prompt_util.py / config_util.py /
     train_lora.py unmodified (a name-only `diffusers` stub is on sys.path; the UNet and
     scheduler objects the reference drives are the oracle's restatements with seeded
     synthetic weights, because diffusers / checkpoints do not exist in this sandbox);
  2. runs train_lora.train() itself on CPU/fp32 for a few iterations (model loading and
     text encoding are the only things patched out, SURVEY.md §2 marks them out of scope);
  3. records per-iteration losses, k, the saved .safetensors keys / checksums;
  4. runs oracle/leco_ref.py on the same seeds and REQUIRES bit-identical results;
  5. writes tests/golden/*.json.

    python tests/golden/make_golden.py
"""
from __future__ import annotations

import contextlib
import io
import json
import os
import sys
import tempfile
import zlib

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import leco_ref  # noqa: E402
from oracle.ref_loader import load_reference  # noqa: E402
from oracle.sched_ref import create_noise_scheduler  # noqa: E402
from oracle.unet_ref import CONFIGS, build_unet  # noqa: E402

GOLDEN_DIR = os.path.dirname(os.path.abspath(__file__))
SEED = 1234


def synthetic_embedding(prompt: str, dim: int) -> torch.Tensor:
    """Stand-in for train_util.encode_prompts (out of scope): N(0,1) [1,77,D] per distinct
    prompt, seeded by the prompt text."""
    g = torch.Generator(device="cpu").manual_seed(zlib.crc32(prompt.encode()) & 0x7FFFFFFF)
    return torch.randn((1, 77, dim), generator=g)


def tensor_digest(t: torch.Tensor) -> dict:
    t = t.detach().to(torch.float64).flatten()
    return {"n": t.numel(), "sum": float(t.sum()), "abs": float(t.abs().sum()),
            "head": [float(x) for x in t[:4]]}


PROMPTS_YAML = """
- target: "van gogh"
  positive: "van gogh"
  unconditional: ""
  neutral: ""
  action: "erase"
  guidance_scale: 1.0
  resolution: 128
  dynamic_resolution: false
  batch_size: 2
- target: "cat"
  positive: "a photo of a cat"
  unconditional: ""
  neutral: "animal"
  action: "enhance"
  guidance_scale: 1.5
  resolution: 128
  batch_size: 1
- target: "oil painting"
  positive: "oil painting"
  unconditional: ""
  neutral: ""
  action: "erase"
  guidance_scale: 1.0
  resolution: 384
  dynamic_resolution: true
  dynamic_crops: true
  batch_size: 1
"""

CONFIG_YAML = """
prompts_file: "{prompts}"
pretrained_model:
  name_or_path: "synthetic:{arch}"
  v2: true
  v_pred: {v_pred}
network:
  type: "lierla"
  rank: 4
  alpha: 1.0
  training_method: "full"
train:
  precision: "float32"
  noise_scheduler: "ddim"
  iterations: {iters}
  lr: 1e-3
  optimizer: "AdamW"
  lr_scheduler: "constant"
  max_denoising_steps: {max_steps}
save:
  name: "golden"
  path: "{out}"
  per_steps: 200
  precision: "float32"
logging:
  use_wandb: false
  verbose: false
other:
  use_xformers: false
"""


def run_reference_train(arch: str, iters: int, max_steps: int, v_pred: bool, unet=None):
    """The reference's UNMODIFIED train_lora.train(config, prompts) on the CPU; `model_util.load_models` is the one
    patched seam and returns `unet` (default: the oracle UNet; tests also pass the engine)."""
    ref = load_reference()
    import importlib
    os.environ.setdefault("WANDB_MODE", "disabled")
    train_lora = importlib.import_module("train_lora")
    assert os.path.abspath(train_lora.__file__).startswith("/root/reference")
    cfg = CONFIGS[arch]
    unet = build_unet(arch, seed=0) if unet is None else unet

    class _Dummy:
        def to(self, *a, **k):
            return self

        def eval(self):
            return self

    def fake_load_models(name, scheduler_name, v2=False, v_pred=False, weight_dtype=torch.float32):
        sched = create_noise_scheduler(scheduler_name, "v_prediction" if v_pred else "epsilon")
        return _Dummy(), _Dummy(), unet, sched

    def fake_encode(tokenizer, text_encoder, prompts):
        return torch.cat([synthetic_embedding(p, cfg.cross_attention_dim) for p in prompts])

    losses, ks = [], []
    orig_loss = ref.prompt_util.PromptEmbedsPair.loss

    def rec_loss(self, **kw):
        out = orig_loss(self, **kw)
        losses.append(float(out.item()))
        return out

    orig_diffusion = ref.train_util.diffusion

    def rec_diffusion(*a, **kw):
        ks.append(int(kw["total_timesteps"]))
        return orig_diffusion(*a, **kw)

    with tempfile.TemporaryDirectory() as tmp:
        pfile = os.path.join(tmp, "prompts.yaml")
        open(pfile, "w").write(PROMPTS_YAML)
        cfile = os.path.join(tmp, "config.yaml")
        open(cfile, "w").write(CONFIG_YAML.format(prompts=pfile, arch=arch, iters=iters,
                                                 max_steps=max_steps, out=os.path.join(tmp, "out"),
                                                 v_pred=str(v_pred).lower()))
        config = ref.config_util.load_config_from_yaml(cfile)
        prompts = ref.prompt_util.load_prompts_from_yaml(config.prompts_file)
        patches = [(ref.model_util, "load_models", fake_load_models),
                   (ref.train_util, "encode_prompts", fake_encode),
                   (ref.train_util, "diffusion", rec_diffusion),
                   (ref.prompt_util.PromptEmbedsPair, "loss", rec_loss),
                   (train_lora, "DEVICE_CUDA", torch.device("cpu"))]
        saved = [(o, n, getattr(o, n)) for o, n, _ in patches]
        try:
            for o, n, v in patches:
                setattr(o, n, v)
            ref.prompt_util.PromptEmbedsCache.prompts.clear()
            torch.manual_seed(SEED)
            with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
                train_lora.train(config, prompts)
        finally:
            for o, n, v in saved:
                setattr(o, n, v)
        from safetensors.torch import load_file
        sd = load_file(os.path.join(tmp, "out", "golden_last.safetensors"))
    # un-patch the oracle unet's Linear forwards for the next user
    return losses, ks, sd, prompts


def run_oracle_train(arch: str, iters: int, max_steps: int, v_pred: bool, prompt_settings):
    cfg = CONFIGS[arch]
    unet = build_unet(arch, seed=0)
    sched = create_noise_scheduler("ddim", "v_prediction" if v_pred else "epsilon")
    torch.manual_seed(SEED)
    with contextlib.redirect_stdout(io.StringIO()):
        net = leco_ref.LoRANetworkRef(unet, rank=4, multiplier=1.0, alpha=1.0, train_method="full")
    opt = torch.optim.AdamW(net.prepare_optimizer_params(), lr=1e-3)
    lrs = torch.optim.lr_scheduler.ConstantLR(opt, factor=1)
    pairs = []
    for s in prompt_settings:
        emb = {p: synthetic_embedding(p, cfg.cross_attention_dim)
               for p in (s.target, s.positive, s.unconditional, s.neutral)}
        pairs.append(leco_ref.PromptPairRef(
            target=emb[s.target], positive=emb[s.positive], unconditional=emb[s.unconditional],
            neutral=emb[s.neutral], guidance_scale=s.guidance_scale, resolution=s.resolution,
            dynamic_resolution=s.dynamic_resolution, batch_size=s.batch_size, action=s.action))
    losses, ks, extra = [], [], []
    for _ in range(iters):
        rec = {}
        losses.append(leco_ref.leco_iteration(unet, sched, net, opt, lrs, pairs,
                                              max_denoising_steps=max_steps, record=rec))
        ks.append(rec["k"])
        extra.append([rec["height"], rec["width"]])
    run_oracle_train.last_hw = extra
    return losses, ks, net.lora_state_dict(torch.float32)


def synthetic_embedding_xl(prompt: str, dim: int, pooled_dim: int):
    g = torch.Generator(device="cpu").manual_seed((zlib.crc32(prompt.encode()) ^ 0x5EED) & 0x7FFFFFFF)
    return torch.randn((1, 77, dim), generator=g), torch.randn((1, pooled_dim), generator=g)


def run_reference_train_xl(arch: str, iters: int, max_steps: int, unet=None):
    """The reference's unmodified train_lora_xl.train() (train_lora_xl.py:40-385) on the oracle SDXL-topology UNet
    (or on `unet`: tests pass the engine)."""
    ref = load_reference()
    import importlib
    os.environ.setdefault("WANDB_MODE", "disabled")
    tx = importlib.import_module("train_lora_xl")
    assert os.path.abspath(tx.__file__).startswith("/root/reference")
    cfg = CONFIGS[arch]
    pooled = cfg.projection_class_embeddings_input_dim - 6 * cfg.addition_time_embed_dim
    unet = build_unet(arch, seed=0) if unet is None else unet

    class _Dummy:
        def to(self, *a, **k):
            return self

        def eval(self):
            return self

        def requires_grad_(self, *a):
            return self

    def fake_load_models_xl(name, scheduler_name, weight_dtype=torch.float32):
        return [_Dummy(), _Dummy()], [_Dummy(), _Dummy()], unet, create_noise_scheduler(scheduler_name, "epsilon")

    def fake_encode_xl(tokenizers, text_encoders, prompts, num_images_per_prompt=1):
        return synthetic_embedding_xl(prompts[0], cfg.cross_attention_dim, pooled)

    losses, ks = [], []
    orig_loss = ref.prompt_util.PromptEmbedsPair.loss

    def rec_loss(self, **kw):
        out = orig_loss(self, **kw)
        losses.append(float(out.item()))
        return out
    orig_diff = ref.train_util.diffusion_xl

    def rec_diff(*a, **kw):
        ks.append(int(kw["total_timesteps"]))
        return orig_diff(*a, **kw)
    with tempfile.TemporaryDirectory() as tmp:
        pfile = os.path.join(tmp, "prompts.yaml")
        # REFERENCE BUG 2: predict_noise_xl (train_util.py:248-251) calls rescale_noise_cfg(noise_pred[2B], noise_pred_text[B])
        # and throws the result away; the shapes only broadcast for B == 1, so the XL loop crashes for batch_size > 1.
        # The XL golden therefore uses batch_size 1 everywhere.
        open(pfile, "w").write(PROMPTS_YAML.replace("batch_size: 2", "batch_size: 1"))
        cfile = os.path.join(tmp, "config.yaml")
        open(cfile, "w").write(CONFIG_YAML.format(prompts=pfile, arch=arch, iters=iters, max_steps=max_steps,
                                                 out=os.path.join(tmp, "out"), v_pred="false"))
        config = ref.config_util.load_config_from_yaml(cfile)
        prompts = ref.prompt_util.load_prompts_from_yaml(config.prompts_file)
        patches = [(ref.model_util, "load_models_xl", fake_load_models_xl),
                   (ref.train_util, "encode_prompts_xl", fake_encode_xl),
                   (ref.train_util, "diffusion_xl", rec_diff),
                   (ref.prompt_util.PromptEmbedsPair, "loss", rec_loss),
                   (tx, "DEVICE_CUDA", torch.device("cpu")),
                   # REFERENCE BUG: train_lora_xl.py:131-138 calls PromptEmbedsXL(<tuple>) but the class unpacks
                   # *args (prompt_util.py:21-23) -> IndexError; as shipped the XL script cannot get past prompt
                   # caching.  The harness unpacks the tuple (the evident intent) so the rest of the loop can be pinned.
                   (tx, "PromptEmbedsXL", lambda t: ref.prompt_util.PromptEmbedsXL(*t))]
        saved = [(o, n, getattr(o, n)) for o, n, _ in patches]
        try:
            for o, n, v in patches:
                setattr(o, n, v)
            ref.prompt_util.PromptEmbedsCache.prompts.clear()
            torch.manual_seed(SEED)
            with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
                tx.train(config, prompts)
        finally:
            for o, n, v in saved:
                setattr(o, n, v)
        from safetensors.torch import load_file
        sd = load_file(os.path.join(tmp, "out", "golden_last.safetensors"))
    return losses, ks, sd, prompts


def run_oracle_train_xl(arch: str, iters: int, max_steps: int, prompt_settings):
    cfg = CONFIGS[arch]
    pooled = cfg.projection_class_embeddings_input_dim - 6 * cfg.addition_time_embed_dim
    unet = build_unet(arch, seed=0)
    sched = create_noise_scheduler("ddim", "epsilon")
    torch.manual_seed(SEED)
    with contextlib.redirect_stdout(io.StringIO()):
        net = leco_ref.LoRANetworkRef(unet, rank=4, multiplier=1.0, alpha=1.0, train_method="full")
    opt = torch.optim.AdamW(net.prepare_optimizer_params(), lr=1e-3)
    lrs = torch.optim.lr_scheduler.ConstantLR(opt, factor=1)
    pairs = []
    for s in prompt_settings:
        emb = {p: leco_ref.EmbedsXL(*synthetic_embedding_xl(p, cfg.cross_attention_dim, pooled))
               for p in (s.target, s.positive, s.unconditional, s.neutral)}
        pairs.append(leco_ref.PromptPairRef(
            target=emb[s.target], positive=emb[s.positive], unconditional=emb[s.unconditional],
            neutral=emb[s.neutral], guidance_scale=s.guidance_scale, resolution=s.resolution,
            dynamic_resolution=s.dynamic_resolution, batch_size=s.batch_size, action=s.action,
            dynamic_crops=s.dynamic_crops))
    losses, ks, extra = [], [], []
    for _ in range(iters):
        rec = {}
        losses.append(leco_ref.leco_iteration_xl(unet, sched, net, opt, lrs, pairs, max_denoising_steps=max_steps,
                                                 record=rec))
        ks.append(rec["k"])
        extra.append({"hw": [rec["height"], rec["width"]], "time_ids": rec["time_ids"]})
    run_oracle_train_xl.last_extra = extra
    return losses, ks, net.lora_state_dict(torch.float32)


def main():
    torch.set_num_threads(4)
    out = {}
    for arch, iters, max_steps, v_pred in (("tiny21", 6, 8, True), ("tiny15", 5, 6, False)):
        ref_losses, ref_ks, ref_sd, prompts = run_reference_train(arch, iters, max_steps, v_pred)
        ora_losses, ora_ks, ora_sd = run_oracle_train(arch, iters, max_steps, v_pred, prompts)
        assert ref_ks == ora_ks, (ref_ks, ora_ks)
        assert ref_losses == ora_losses, (ref_losses, ora_losses)
        assert sorted(ref_sd.keys()) == sorted(ora_sd.keys())  # safetensors re-orders keys on disk
        for k in ref_sd:
            assert torch.equal(ref_sd[k], ora_sd[k]), k
        keys = sorted(ref_sd.keys())
        out[arch] = {
            "seed": SEED, "iters": iters, "max_denoising_steps": max_steps, "v_pred": v_pred,
            "lr": 1e-3, "losses": ref_losses, "k": ref_ks, "n_keys": len(keys),
            "first_keys": keys[:6], "last_keys": keys[-3:],
            "digests": {k: tensor_digest(ref_sd[k]) for k in keys[:12] + keys[-12:]},
            "total_abs": float(sum(v.double().abs().sum() for v in ref_sd.values())),
            "hw": run_oracle_train.last_hw,   # (height, width) per iteration: the dynamic_resolution prompt draws buckets
        }
        print(arch, "hw", out[arch]["hw"], "k", ref_ks)
        print(arch, "reference == oracle bit-exact;", "losses", ref_losses, "k", ref_ks)
    assert any(hw != [128, 128] for a in ("tiny21", "tiny15") for hw in out[a]["hw"]), \
        "the dynamic_resolution prompt was never drawn"
    ref_losses, ref_ks, ref_sd, prompts = run_reference_train_xl("tinyxl", 5, 6)
    ora_losses, ora_ks, ora_sd = run_oracle_train_xl("tinyxl", 5, 6, prompts)
    assert ref_ks == ora_ks and ref_losses == ora_losses, (ref_ks, ora_ks, ref_losses, ora_losses)
    assert sorted(ref_sd.keys()) == sorted(ora_sd.keys())
    for k in ref_sd:
        assert torch.equal(ref_sd[k], ora_sd[k]), k
    keys = sorted(ref_sd.keys())
    out["tinyxl"] = {"seed": SEED, "iters": 5, "max_denoising_steps": 6, "lr": 1e-3, "losses": ref_losses,
                     "k": ref_ks, "n_keys": len(keys),
                     "total_abs": float(sum(v.double().abs().sum() for v in ref_sd.values())),
                     "extra": run_oracle_train_xl.last_extra}   # (h, w) and add_time_ids per iteration (dynamic crops)
    assert any(e["time_ids"][2:4] != [0.0, 0.0] for e in out["tinyxl"]["extra"]), "dynamic crops never drawn"
    print("tinyxl (train_lora_xl.train) reference == oracle bit-exact;", "losses", ref_losses, "k", ref_ks)
    out["_meta"] = {"torch": torch.__version__,
                    "how": "reference train_lora.train() (unmodified) on oracle UNet/DDIM, CPU fp32"}
    with open(os.path.join(GOLDEN_DIR, "leco_train_golden.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", os.path.join(GOLDEN_DIR, "leco_train_golden.json"))


if __name__ == "__main__":
    main()
