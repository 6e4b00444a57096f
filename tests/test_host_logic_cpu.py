"""CPU: host-side mirrors of the reference interface (scheduler, LoRANetwork discovery / naming /
init / export) against the oracle, and against the live reference when it is present."""
import contextlib
import io
import os

import pytest
import torch

from leco_b200 import lora as plora
from leco_b200.scheduler import DDIMScheduler, create_noise_scheduler
from leco_b200.unet import SPECS, EngineUNet
from oracle import leco_ref
from oracle.ref_loader import reference_available
from oracle.sched_ref import create_noise_scheduler as oracle_sched
from oracle.unet_ref import build_unet
from tests import torch_backend


@pytest.mark.parametrize("ptype", ["epsilon", "v_prediction"])
def test_ddim_matches_oracle(ptype):
    a, b = DDIMScheduler(ptype), oracle_sched("ddim", ptype)
    for n in (50, 30, 1000):
        a.set_timesteps(n)
        b.set_timesteps(n)
        assert a.timesteps.tolist() == b.timesteps.tolist()
    a.set_timesteps(50)
    b.set_timesteps(50)
    g = torch.Generator().manual_seed(0)
    x, e = torch.randn(2, 4, 8, 8, generator=g), torch.randn(2, 4, 8, 8, generator=g)
    for t in (980, 500, 20, 0):
        cx, ce = a.coefficients(t)              # the affine map the GPU kernels apply (guided_step / axpby)
        pa = cx * x + ce * e
        pb = b.step(e, t, x).prev_sample
        assert (pa - pb).abs().max().item() < 2e-6
    with pytest.raises(RuntimeError):           # product scheduler has no host arithmetic path
        a.step(e, 980, x)
    assert a.init_noise_sigma == 1.0 and a.scale_model_input(x, 3) is x
    with pytest.raises(ValueError):
        create_noise_scheduler("nope")


def _nets(arch="tiny21"):
    eng = EngineUNet(SPECS[arch], backend=torch_backend)
    ora = build_unet(arch)
    torch.manual_seed(21)
    with contextlib.redirect_stdout(io.StringIO()):
        a = plora.LoRANetwork(eng, rank=4, multiplier=1.0, alpha=1.0)
    torch.manual_seed(21)
    with contextlib.redirect_stdout(io.StringIO()):
        b = leco_ref.LoRANetworkRef(ora, rank=4, multiplier=1.0, alpha=1.0)
    return a, b


def test_lora_network_mirror_matches_oracle():
    a, b = _nets()
    assert len(a.unet_loras) == 192
    assert list(a.state_dict().keys()) == list(b.state_dict().keys())
    for (k, va), vb in zip(a.state_dict().items(), b.state_dict().values()):
        assert torch.equal(va, vb), k            # same RNG draws in the same order
    assert a.unet_loras[0].scale == 0.25 and float(a.unet_loras[0].alpha) == 1.0
    pa, pb = a.prepare_optimizer_params(), b.prepare_optimizer_params()
    assert len(pa) == 1 and len(pa[0]["params"]) == len(pb[0]["params"]) == 384
    assert all(l.multiplier == 1.0 for l in a.unet_loras)
    with a:
        pass
    assert all(l.multiplier == 0 for l in a.unet_loras)
    with pytest.raises(NotImplementedError):
        plora.LoRANetwork(EngineUNet(SPECS["tiny21"], backend=torch_backend), train_method="bogus")


def test_save_weights_format(tmp_path):
    from safetensors.torch import load_file
    a, b = _nets()
    f = str(tmp_path / "x_last.safetensors")
    a.save_weights(f, dtype=torch.bfloat16)
    sd = load_file(f)
    want = b.lora_state_dict(torch.bfloat16)
    assert sorted(sd.keys()) == sorted(want.keys())
    for k in sd:
        assert sd[k].dtype == torch.bfloat16 and torch.equal(sd[k], want[k]), k
    assert "lora_unet_down_blocks_0_attentions_0_proj_in.alpha" in sd
    assert sd["lora_unet_mid_block_attentions_0_transformer_blocks_0_attn2_to_k.lora_down.weight"].shape == (4, 128)


def test_rng_helpers_draw_like_the_oracle():
    """Bucket resolution and SDXL add_time_ids (with dynamic crops) consume the global CPU generator exactly like
    the pinned oracle (train_util.py:295-330, :404-416): same values, same generator state afterwards."""
    from leco_b200 import train_util as tu
    from leco_b200.trainer import get_random_resolution_in_bucket as bucket
    from oracle import leco_ref
    for seed in (0, 1, 7):
        torch.manual_seed(seed)
        a = [bucket(512), bucket(384), tu.get_add_time_ids(192, 320, dynamic_crops=True),
             tu.get_add_time_ids(128, 128, dynamic_crops=False), torch.rand(1)]
        torch.manual_seed(seed)
        b = [leco_ref.get_random_resolution_in_bucket(512), leco_ref.get_random_resolution_in_bucket(384),
             leco_ref.get_add_time_ids(192, 320, dynamic_crops=True),
             leco_ref.get_add_time_ids(128, 128, dynamic_crops=False), torch.rand(1)]
        assert a[0] == b[0] and a[1] == b[1]
        assert torch.equal(a[2], b[2]) and torch.equal(a[3], b[3]) and torch.equal(a[4], b[4])
    assert tu.get_add_time_ids(1024, 1024).tolist() == [[1024, 1024, 0, 0, 1024, 1024]]   # SURVEY 8c known answer (10)


def test_trainer_prompt_dedup_rules():
    """LecoTrainer batches identical prompts of the LoRA-off passes; identity must include the pooled embedding."""
    from leco_b200.trainer import EmbedsXL, LecoTrainer, PromptPair
    g = torch.Generator().manual_seed(0)
    t1, t2 = torch.randn(1, 77, 8, generator=g), torch.randn(1, 77, 8, generator=g)
    p1, p2 = torch.randn(1, 4, generator=g), torch.randn(1, 4, generator=g)
    same = LecoTrainer._same_prompt
    assert same(t1, t1) and same(t1, t1.clone()) and not same(t1, t2)
    assert same(EmbedsXL(t1, p1), EmbedsXL(t1.clone(), p1.clone()))
    assert not same(EmbedsXL(t1, p1), EmbedsXL(t1, p2)) and not same(EmbedsXL(t1, p1), t1)
    assert PromptPair(t1, t1, t2, t2, guidance_scale=1.5, action="erase").signed_guidance() == -1.5
    assert PromptPair(t1, t1, t2, t2, guidance_scale=1.5, action="enhance").signed_guidance() == 1.5
    with pytest.raises(ValueError):
        PromptPair(t1, t1, t2, t2, action="nope").signed_guidance()


def test_load_weights_round_trip(tmp_path):
    """save_weights -> load_weights restores every adapter tensor in place."""
    a, _ = _nets()
    g = torch.Generator().manual_seed(3)
    for l in a.unet_loras:
        l.lora_up.weight.data = 0.1 * torch.randn(l.lora_up.weight.shape, generator=g)
    want = {k: v.clone() for k, v in a.state_dict().items() if k.startswith("lora")}
    f = str(tmp_path / "ckpt.safetensors")
    a.save_weights(f)
    ptrs = [l.lora_up.weight.data_ptr() for l in a.unet_loras]
    for l in a.unet_loras:
        l.lora_up.weight.data.zero_()
        l.lora_down.weight.data.zero_()
    missing, unexpected = a.load_weights(f)
    assert not missing and not unexpected
    assert ptrs == [l.lora_up.weight.data_ptr() for l in a.unet_loras]      # in place: flat buffer / graphs stay valid
    for k, v in a.state_dict().items():
        if k.startswith("lora"):
            assert torch.equal(v, want[k]), k
    import pytest
    sd = {k: v for k, v in want.items() if "mid_block" not in k}
    torch.save(sd, str(tmp_path / "partial.pt"))
    with pytest.raises(KeyError):
        a.load_weights(str(tmp_path / "partial.pt"))
    missing, _ = a.load_weights(str(tmp_path / "partial.pt"), strict=False)
    assert missing and all("mid_block" in k for k in missing)


def test_c3lier_aliasing_contract():
    """extending DEFAULT_TARGET_REPLACE in place (train_lora.py:44-46) adds the conv blocks."""
    eng = EngineUNet(SPECS["tiny15"], backend=torch_backend)
    saved = list(plora.DEFAULT_TARGET_REPLACE)
    try:
        plora.DEFAULT_TARGET_REPLACE += plora.UNET_TARGET_REPLACE_MODULE_CONV
        with contextlib.redirect_stdout(io.StringIO()):
            net = plora.LoRANetwork(eng, rank=8, alpha=1.0)
        assert len(net.unet_loras) == 278
    finally:
        del plora.DEFAULT_TARGET_REPLACE[len(saved):]


@pytest.mark.skipif(not reference_available(), reason="reference sources only exist in the build container")
def test_reference_lora_network_patches_engine_tree():
    """The reference's OWN LoRANetwork must find the same targets on the engine's holder tree."""
    from oracle.ref_loader import load_reference
    ref = load_reference()
    for arch, n in (("tiny21", 192), ("tiny15", 192), ("tinyxl", 302)):
        eng = EngineUNet(SPECS[arch], backend=torch_backend)
        ora = build_unet(arch)
        with contextlib.redirect_stdout(io.StringIO()):
            a = ref.lora.LoRANetwork(eng, rank=4, multiplier=1.0, alpha=1.0)
            b = ref.lora.LoRANetwork(ora, rank=4, multiplier=1.0, alpha=1.0)
        assert len(a.unet_loras) == n
        assert [l.lora_name for l in a.unet_loras] == [l.lora_name for l in b.unet_loras]
        from leco_b200.unet import find_adapter
        sites = [s for s in (eng.pack("cpu", torch.float32) or eng).lora_sites() if s.adapters() is not None]
        assert sum(len(s.members) for s in sites) == n


@pytest.mark.skipif(not reference_available(), reason="reference sources only exist in the build container")
@pytest.mark.parametrize("arch,c3lier,rank", [("tiny21", False, 4), ("tiny15", True, 8)])
def test_saved_file_is_byte_identical_to_the_reference_export(tmp_path, arch, c3lier, rank):
    """north_star: '.safetensors export stays bit-compatible'.  The reference's OWN LoRANetwork.save_weights (lora.py:212-229)
    on the oracle UNet and the mirror's on the engine tree, same seed, same (trained-looking) weights: the two files are
    the same BYTES (header order, dtypes, shapes, tensor data), for lierla and for c3lier (3x3 conv adapters)."""
    from oracle.ref_loader import load_reference
    ref = load_reference()
    eng, ora = EngineUNet(SPECS[arch], backend=torch_backend), build_unet(arch)
    saved_m, saved_r = list(plora.DEFAULT_TARGET_REPLACE), list(ref.lora.DEFAULT_TARGET_REPLACE)
    try:
        if c3lier:                                   # train_lora.py:44-46
            plora.DEFAULT_TARGET_REPLACE += plora.UNET_TARGET_REPLACE_MODULE_CONV
            ref.lora.DEFAULT_TARGET_REPLACE += ref.lora.UNET_TARGET_REPLACE_MODULE_CONV
        with contextlib.redirect_stdout(io.StringIO()):
            torch.manual_seed(5)
            a = plora.LoRANetwork(eng, rank=rank, multiplier=1.0, alpha=1.0)
            torch.manual_seed(5)
            b = ref.lora.LoRANetwork(ora, rank=rank, multiplier=1.0, alpha=1.0)
    finally:
        plora.DEFAULT_TARGET_REPLACE[:] = saved_m
        ref.lora.DEFAULT_TARGET_REPLACE[:] = saved_r
    g = torch.Generator().manual_seed(9)
    for la, lb in zip(a.unet_loras, b.unet_loras):   # lora_up is zero at init: give both the same non-trivial values
        w = 0.1 * torch.randn(la.lora_up.weight.shape, generator=g)
        la.lora_up.weight.data.copy_(w)
        lb.lora_up.weight.data.copy_(w)
    for dtype in (torch.bfloat16, None):
        fa, fb = str(tmp_path / f"a_{dtype}.safetensors"), str(tmp_path / f"b_{dtype}.safetensors")
        a.save_weights(fa, dtype=dtype)
        b.save_weights(fb, dtype=dtype)
        assert open(fa, "rb").read() == open(fb, "rb").read(), (arch, dtype)
    # the metadata argument reaches the file header unchanged too
    a.save_weights(str(tmp_path / "ma.safetensors"), dtype=torch.float16, metadata={"ss_network_dim": str(rank)})
    b.save_weights(str(tmp_path / "mb.safetensors"), dtype=torch.float16, metadata={"ss_network_dim": str(rank)})
    assert open(tmp_path / "ma.safetensors", "rb").read() == open(tmp_path / "mb.safetensors", "rb").read()


def test_lr_scheduler_factory_matches_torch():
    """train_util.get_lr_scheduler (train_util.py:373-401) drives the fused optimizer's lr through the same torch
    scheduler classes and arguments; the reference steps it once per iteration (train_lora.py:281)."""
    from leco_b200 import train_util as tu

    class Opt:   # stands for FlatOptimizer: only `.lr` is touched
        lr = 1e-3

    for name in ("constant", "cosine", "cosine_with_restarts", "step", "linear"):
        o = Opt()
        sched = tu.get_lr_scheduler(name, o, max_iterations=200, lr_min=1e-5)
        ref_opt = torch.optim.SGD([torch.nn.Parameter(torch.zeros(1))], lr=1e-3)
        S = torch.optim.lr_scheduler
        ref = {"constant": lambda: S.ConstantLR(ref_opt, factor=1),
               "cosine": lambda: S.CosineAnnealingLR(ref_opt, T_max=200, eta_min=1e-5),
               "cosine_with_restarts": lambda: S.CosineAnnealingWarmRestarts(ref_opt, T_0=20, T_mult=2, eta_min=1e-5),
               "step": lambda: S.StepLR(ref_opt, step_size=2, gamma=0.999),
               "linear": lambda: S.LinearLR(ref_opt, start_factor=0.5, total_iters=2)}[name]()
        for _ in range(50):
            assert o.lr == ref_opt.param_groups[0]["lr"], name
            sched.step()
            ref_opt.step()
            ref.step()
    with pytest.raises(ValueError):
        tu.get_lr_scheduler("bogus", Opt(), 10, 1e-5)


def test_optimizer_factory_error_behaviour():
    """train_util.get_optimizer (train_util.py:333-370): same names, same ValueError texts for unknown names."""
    from leco_b200 import train_util as tu
    for name in ("adamw", "AdamW", "adam", "lion"):
        assert callable(tu.get_optimizer(name))
    with pytest.raises(ValueError, match="Optimizer must be adam, adamw, lion or Prodigy"):
        tu.get_optimizer("sgd")
    with pytest.raises(ValueError, match="DAdapt optimizer must be"):
        tu.get_optimizer("dadaptsgd")
    with pytest.raises(ValueError, match="8bit optimizer must be"):
        tu.get_optimizer("sgd8bit")
    for name in ("dadaptadam", "adam8bit", "prodigy"):   # optional third-party packages, absent here as in the reference image
        with pytest.raises(NotImplementedError):
            tu.get_optimizer(name)


@pytest.mark.parametrize("name", ["ddim", "ddpm", "lms", "euler_a"])
@pytest.mark.parametrize("ptype", ["epsilon", "v_prediction"])
def test_scheduler_coefficient_rows_match_oracle(name, ptype):
    """model_util.create_noise_scheduler's four schedulers (model_util.py:230-278): the product states every update as
    x' = cx x + ce m + cn noise + sum_j l_j d_{-j} (d = dx x + dg m); applied in plain fp64 torch those rows must
    reproduce the oracle's restatement of diffusers' step(), its timesteps, input scaling and init_noise_sigma."""
    prod, ora = create_noise_scheduler(name, ptype), oracle_sched(name, ptype)
    assert abs(float(prod.init_noise_sigma) - float(ora.init_noise_sigma)) < 1e-5
    for n in (50, 1000):
        prod.set_timesteps(n)
        ora.set_timesteps(n)
        assert torch.allclose(prod.timesteps.double(), ora.timesteps.double(), atol=1e-9)
    # the 1000-step grid of the four predictions (train_lora.py:195-199)
    for t in (999, 499, 19):
        x = torch.randn(2, 4, 8, 8, dtype=torch.float64)
        assert torch.allclose(ora.scale_model_input(x, ora.timesteps[999 - t]), x * prod.in_scale_at_train_timestep(t),
                              rtol=1e-5, atol=1e-7)
    prod.set_timesteps(50)
    ora.set_timesteps(50)
    g = torch.Generator().manual_seed(3)
    x = torch.randn(2, 4, 8, 8, generator=g, dtype=torch.float64) * float(ora.init_noise_sigma)
    hist = []
    for i, t in enumerate(ora.timesteps[:9]):
        m = torch.randn(x.shape, generator=g, dtype=torch.float64)
        noise = torch.randn(x.shape, generator=g, dtype=torch.float64)
        row = prod.plan(i)
        _, cx, ce, cn, s_in, dx, dg, l0, l1, l2, l3, slot = row
        assert torch.allclose(ora.scale_model_input(x, t), x * s_in, rtol=1e-5, atol=1e-7)
        kw = {"noise": noise} if name in ("ddpm", "euler_a") else {}
        ref = ora.step(m, t, x, **kw).prev_sample
        got = cx * x + ce * m + cn * noise
        if name == "lms":
            hist.append(dx * x + dg * m)
            hist = hist[-4:]
            assert int(slot) == i % 4
            for c, h in zip((l0, l1, l2, l3), reversed(hist)):
                got = got + c * h
        else:
            assert (dx, dg, l0, l1, l2, l3) == (0.0,) * 6
        assert torch.allclose(got, ref.double(), rtol=2e-4, atol=2e-4 * float(ref.abs().max())), (name, ptype, i)
        x = ref.double()
    with pytest.raises(ValueError):
        create_noise_scheduler("heun")


@pytest.mark.skipif(not reference_available(), reason="needs /root/reference")
def test_config_mirror_parses_the_reference_examples_like_the_reference():
    """leco_b200.config_util (used where the reference's files are absent) vs the reference's own config_util /
    prompt_util on every example YAML: same values field by field (incl. the `lr: 1e-4` string coercion)."""
    import glob
    import sys
    from leco_b200 import config_util as cu
    from oracle.ref_loader import load_reference
    ref = load_reference()
    for path in sorted(glob.glob("/root/reference/examples/*config*.yaml")):
        mine, theirs = cu.load_config_from_yaml(path), ref.config_util.load_config_from_yaml(path)
        for section in ("pretrained_model", "network", "train", "save", "logging", "other"):
            a, b = getattr(mine, section), getattr(theirs, section)
            for k, v in b.dict().items():
                assert getattr(a, k) == v, (path, section, k)
        assert mine.prompts_file == theirs.prompts_file
    for path in sorted(glob.glob("/root/reference/examples/*prompts*.yaml")):
        mine, theirs = cu.load_prompts_from_yaml(path), ref.prompt_util.load_prompts_from_yaml(path)
        assert len(mine) == len(theirs)
        for a, b in zip(mine, theirs):
            for k, v in b.dict().items():
                assert getattr(a, k) == v, (path, k)


def test_shipped_example_config_parses_and_names_the_baseline_workload():
    """examples/sd21_erase.yaml is BASELINE configs[1] in the driver's own format: every key is one the mirror (and the
    reference's RootConfig) defines, and it selects the workload bench.py measures."""
    from leco_b200 import config_util, train_util
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    c = config_util.load_config_from_yaml(os.path.join(root, "examples", "sd21_erase.yaml"))
    ps = config_util.load_prompts_from_yaml(os.path.join(root, "examples", "erase_prompts.yaml"))
    assert (c.pretrained_model.name_or_path, c.pretrained_model.v2, c.pretrained_model.v_pred) == ("sd21", True, True)
    assert (c.network.type, c.network.rank, c.train.precision, c.train.max_denoising_steps) == ("lierla", 4, "bfloat16", 50)
    assert config_util.parse_precision(c.train.precision) == torch.bfloat16 and c.train.lr == 1e-4
    train_util.get_optimizer(c.train.optimizer)               # a known optimizer name
    assert (ps[0].target, ps[0].action, ps[0].batch_size, ps[0].resolution) == ("van gogh", "erase", 2, 512)
    import bench
    b = bench.CONFIGS["sd21"]
    assert (b["arch"], b["rank"], b["batch"], b["res"]) == ("sd21", c.network.rank, ps[0].batch_size, ps[0].resolution)


@pytest.mark.skipif(not reference_available(), reason="reference sources only exist in the build container")
def test_train_util_surface_is_complete_and_the_small_helpers_equal_the_reference():
    """Every function the reference's train_util.py defines exists under the same name in leco_b200.train_util, and the
    ones that are plain arithmetic / RNG draws (apply_noise_offset :36-40, get_initial_latents :43-57, concat_embeddings
    :133-138, rescale_noise_cfg :196-214, get_add_time_ids :295-330, get_random_resolution_in_bucket :404-416) return
    the reference's values bit for bit and leave the global generator in the same state."""
    import inspect
    import re
    from types import SimpleNamespace as NS
    from leco_b200 import train_util as tu
    from oracle.ref_loader import load_reference
    ref = load_reference().train_util
    names = set(re.findall(r"^def (\w+)", open(inspect.getsourcefile(ref)).read(), re.M))
    assert names and not [n for n in names if not callable(getattr(tu, n, None))]
    sched = NS(init_noise_sigma=14.6)
    g = torch.Generator().manual_seed(3)
    a, b = torch.randn(2, 4, 8, 8, generator=g), torch.randn(2, 4, 8, 8, generator=g)
    u, c = torch.randn(1, 77, 16, generator=g), torch.randn(1, 77, 16, generator=g)
    calls = [lambda m: m.apply_noise_offset(a, 0.1), lambda m: m.get_initial_latents(sched, 2, 256, 320, 3),
             lambda m: m.get_random_noise(3, 128, 192), lambda m: m.concat_embeddings(u, c, 3),
             lambda m: m.rescale_noise_cfg(a, b, guidance_rescale=0.7), lambda m: m.get_add_time_ids(320, 256, dynamic_crops=True),
             lambda m: torch.tensor(m.get_random_resolution_in_bucket(512)), lambda m: torch.tensor(m.get_random_resolution_in_bucket(384))]
    for i, f in enumerate(calls):
        torch.manual_seed(100 + i)
        want, state_want = f(ref), torch.get_rng_state()
        torch.manual_seed(100 + i)
        got, state_got = f(tu), torch.get_rng_state()
        assert got.shape == want.shape and got.dtype == want.dtype and torch.equal(got, want), i
        assert torch.equal(state_got, state_want), i


@pytest.mark.skipif(not reference_available(), reason="reference sources only exist in the build container")
def test_prompt_util_mirror_against_the_live_reference():
    """leco_b200.prompt_util carries the reference's names: PromptEmbedsPair(loss_fn, target, positive, unconditional,
    neutral, settings).loss(**kwargs) gives the reference's erase / enhance objective bit for bit (prompt_util.py:107-148),
    is a trainer.PromptPair (signed guidance, settings fields), PromptEmbedsXL is positional, PromptEmbedsCache returns
    None for unknown prompts and shares its dict across instances like the reference's."""
    from leco_b200 import prompt_util as pu
    from leco_b200.trainer import EmbedsXL, PromptPair
    from oracle.ref_loader import load_reference
    ref = load_reference().prompt_util
    g = torch.Generator().manual_seed(0)
    t, p, u, n = (torch.randn(2, 4, 8, 8, generator=g) for _ in range(4))
    emb = torch.randn(1, 77, 16, generator=g)
    crit = torch.nn.MSELoss()
    for action, gs in (("erase", 1.0), ("enhance", 1.5), ("erase", 3.0)):
        kw = dict(target="a", positive="b", unconditional="", neutral="c", action=action, guidance_scale=gs,
                  resolution=384, dynamic_resolution=True, batch_size=3, dynamic_crops=True)
        a = pu.PromptEmbedsPair(crit, emb, emb, emb, emb, pu.PromptSettings(**kw))
        b = ref.PromptEmbedsPair(crit, emb, emb, emb, emb, ref.PromptSettings(**kw))
        la = a.loss(target_latents=t, positive_latents=p, unconditional_latents=u, neutral_latents=n)
        lb = b.loss(target_latents=t, positive_latents=p, unconditional_latents=u, neutral_latents=n)
        assert torch.equal(la, lb)
        assert isinstance(a, PromptPair) and a.signed_guidance() == (-gs if action == "erase" else gs)
        for f in ("guidance_scale", "resolution", "dynamic_resolution", "batch_size", "dynamic_crops", "action"):
            assert getattr(a, f) == getattr(b, f), f
    a.action = "nope"
    with pytest.raises(ValueError, match="action must be erase or enhance"):
        a.loss(target_latents=t, positive_latents=p, unconditional_latents=u, neutral_latents=n)
    xl = pu.PromptEmbedsXL(emb, emb[:, 0])
    assert isinstance(xl, EmbedsXL) and xl.text_embeds is emb and xl.pooled_embeds.shape == (1, 16)
    c1, c2 = pu.PromptEmbedsCache(), pu.PromptEmbedsCache()
    c1["van gogh"] = emb
    assert c2["van gogh"] is emb and c1["unknown"] is None          # class-level dict, as in prompt_util.py:31
    assert set(pu.ACTION_TYPES) == {"erase", "enhance"}
    from leco_b200 import lora as plora2, model_util as mu, train_lora as tl
    assert plora2.LORA_PREFIX_UNET == "lora_unet" and tl.NUM_IMAGES_PER_PROMPT == 1 and callable(tl.flush)
    assert set(mu.AVAILABLE_SCHEDULERS) == {"ddim", "ddpm", "lms", "euler_a"} and mu.DIFFUSERS_CACHE_DIR is None
    for name in ("load_diffusers_model", "load_checkpoint_model", "load_diffusers_model_xl", "load_checkpoint_model_xl"):
        with pytest.raises(FileNotFoundError):
            getattr(mu, name)("/nonexistent/model")


def test_optimizer_args_parsing_and_save_cadence():
    """train_lora.py:81-87 ("k=v k=v" through ast.literal_eval) and :292-309 (periodic saves skip i == 0 and the last)."""
    from leco_b200.train_lora import parse_optimizer_args
    assert parse_optimizer_args("") == {} and parse_optimizer_args(None) == {}
    assert parse_optimizer_args("weight_decay=0.1 betas=(0.9,0.99)") == {"weight_decay": 0.1, "betas": (0.9, 0.99)}
    saves = [i for i in range(500) if i % 200 == 0 and i != 0 and i != 499]
    assert saves == [200, 400]


def test_run_logger_mirrors_the_reference_logging(capsys):
    """train_lora.py:38-52, 274-277: metadata = {"prompts": joined .json(), "config": .json()}, printed when verbose;
    wandb.init(project="LECO_<name>", config=metadata) and wandb.log({"loss", "iteration", "lr"}) per iteration; only
    rank 0 logs, and the loop reads the loss back only when something consumes it."""
    from types import SimpleNamespace as NS
    from leco_b200 import config_util
    from leco_b200.train_lora import RunLogger

    class FakeWandb:
        def __init__(self):
            self.calls = []

        def init(self, **kw):
            self.calls.append(("init", kw))

        def log(self, d):
            self.calls.append(("log", d))

        def finish(self):
            self.calls.append(("finish", None))

    def cfg(verbose, use_wandb):
        return NS(logging=NS(verbose=verbose, use_wandb=use_wandb), save=NS(name="van_gogh"), json=lambda: '{"c": 1}')
    prompts = [NS(json=lambda: '{"target": "van gogh"}'), NS(json=lambda: '{"target": "cat"}')]
    quiet = RunLogger(cfg(False, False), prompts)
    assert not quiet.needs_loss and capsys.readouterr().out == ""
    w = FakeWandb()
    log = RunLogger(cfg(True, True), prompts, rank=0, wandb_module=w)
    assert log.needs_loss and log.metadata == {"prompts": '{"target": "van gogh"},{"target": "cat"}', "config": '{"c": 1}'}
    assert w.calls == [("init", {"project": "LECO_van_gogh", "config": log.metadata})]
    assert str(log.metadata) in capsys.readouterr().out
    log.iteration(3, 0.25, 1e-4, 17)
    log.finish()
    assert w.calls[1:] == [("log", {"loss": 0.25, "iteration": 3, "lr": 1e-4}), ("finish", None)]
    assert "iteration 3" in capsys.readouterr().out
    w2 = FakeWandb()
    other = RunLogger(cfg(True, True), prompts, rank=1, wandb_module=w2)      # data parallel: ranks > 0 stay silent
    assert not other.needs_loss and w2.calls == [] and capsys.readouterr().out == ""
    # the mirror's own config / prompt objects serialise too
    if reference_available():
        import json
        from leco_b200.train_lora import _json
        root = config_util.load_config_from_yaml("/root/reference/examples/config.yaml")
        ps = config_util.load_prompts_from_yaml("/root/reference/examples/prompts.yaml")
        assert json.loads(_json(root))["train"]["iterations"] == root.train.iterations
        assert json.loads(_json(ps[0]))["target"] == ps[0].target
        assert RunLogger(root, ps).metadata["config"] == _json(root)


def test_attention_dispatch_rule(monkeypatch):
    """Which attention path a call takes (leco_b200/ops.py::attention): the fused forward covers head dims <= 192 on
    every no-grad pass (SD1.5's 80 / 160 included), the fused backward head dims <= 64 only, and the deterministic
    switch keeps the materialised path on the grad pass (the fused backward reduces dQ with fp32 atomics)."""
    from leco_b200 import ops
    calls = []
    monkeypatch.setattr(ops, "flash_attention", lambda *a, **k: calls.append("flash") or "o")
    monkeypatch.setattr(ops, "flash_attention_lse", lambda *a, **k: calls.append("flash_lse") or ("o", "lse"))
    monkeypatch.setattr(ops, "attention_v0", lambda *a, **k: calls.append("v0") or ("o", None))
    monkeypatch.setattr(ops, "ATTENTION_IMPL", "flash")
    monkeypatch.setattr(ops, "ATTENTION_BWD_IMPL", "flash")
    monkeypatch.setattr(ops, "FLASH_WIDE_MAX_D", 192)
    monkeypatch.setattr(ops, "FLASH_V_MODE", 0)
    monkeypatch.setattr(ops, "_DETERMINISTIC", [False])

    def path(d, grad):
        calls.clear()
        ops.attention(None, None, None, 2, 64, 64, 8, d, d ** -0.5, grad)
        return calls[-1]

    assert [path(d, False) for d in (40, 64, 80, 160, 192)] == ["flash"] * 5
    assert path(200, False) == "v0" and path(44, False) == "v0"          # too wide / not a multiple of 8
    assert path(64, True) == "flash_lse" and path(40, True) == "flash_lse"
    assert path(80, True) == "v0" and path(160, True) == "v0"            # grad pass: fused backward covers d <= 64
    monkeypatch.setattr(ops, "_DETERMINISTIC", [True])
    assert path(64, True) == "v0" and path(64, False) == "flash"
    monkeypatch.setattr(ops, "_DETERMINISTIC", [False])
    monkeypatch.setattr(ops, "FLASH_WIDE_MAX_D", 64)                     # LECO_FLASH_WIDE=0
    assert path(80, False) == "v0" and path(64, False) == "flash"
    monkeypatch.setattr(ops, "FLASH_WIDE_MAX_D", 192)
    monkeypatch.setattr(ops, "FLASH_V_MODE", 1)                          # the wide kernel takes V in place only
    assert path(80, False) == "v0" and path(64, False) == "flash"


def test_timeline_attribution_by_end_times():
    """tests/gpu_checks/timeline_step.py::table charges a kernel with the time by which it extends the timeline (a
    PDL kernel starts while its predecessor still runs), and reports time with no kernel resident as idle."""
    from tests.gpu_checks.timeline_step import table
    ev = [{"name": "a", "ts": 0.0, "dur": 10.0},      # 0..10
          {"name": "b", "ts": 2.0, "dur": 11.0},      # starts under a (PDL), ends at 13: extends the timeline by 3
          {"name": "c", "ts": 4.0, "dur": 5.0},       # entirely hidden: 0
          {"name": "a", "ts": 20.0, "dur": 5.0}]      # 7 us with nothing resident, then 5
    md = table(ev, "t")
    rows = {l.split("|")[1].strip(" `"): [c.strip() for c in l.split("|")[2:]] for l in md.splitlines() if l.startswith("| `")}
    assert "span: 0.025 ms" in md and "idle (no kernel resident): 0.007 ms" in md and "busy: 0.018 ms" in md
    assert rows["a"][0] == "2" and float(rows["a"][1]) == 0.015 and rows["b"][1] == "0.003" and rows["c"][1] == "0.000"
