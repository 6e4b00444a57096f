"""CPU: the engine's forward program and tape backward (leco_b200/unet.py) driven by the
plain-torch test double must reproduce the oracle UNet (forward) and the oracle's autograd
LoRA gradients, in fp32.  This validates wiring/algebra only; kernels are checked on the GPU."""
import contextlib
import io

import pytest
import torch

from leco_b200.unet import SPECS, EngineUNet
from oracle import leco_ref
from oracle.unet_ref import CONFIGS, build_unet
from tests import torch_backend


def _inputs(arch, n=2, hw=16, seed=3):
    cfg = CONFIGS[arch]
    g = torch.Generator().manual_seed(seed)
    x = torch.randn((n, 4, hw, hw), generator=g)
    ctx = torch.randn((n, 77, cfg.cross_attention_dim), generator=g)
    added = None
    if cfg.addition_embed_type:
        pooled = cfg.projection_class_embeddings_input_dim - 6 * cfg.addition_time_embed_dim
        added = {"text_embeds": torch.randn((n, pooled), generator=g),
                 "time_ids": torch.tensor([[128., 128, 0, 0, 128, 128]] * n)}
    return x, ctx, added


def _engine(arch, oracle):
    eng = EngineUNet(SPECS[arch], backend=torch_backend)
    eng.load_state_dict(oracle.state_dict())
    eng._act_dtype = torch.float32
    eng.requires_grad_(False)
    return eng


@pytest.mark.parametrize("arch", ["tiny21", "tiny15", "tinyxl"])
def test_forward_matches_oracle(arch):
    torch.manual_seed(0)
    oracle = build_unet(arch)
    eng = _engine(arch, oracle)
    x, ctx, added = _inputs(arch)
    t = torch.tensor(481)
    with torch.no_grad():
        ref = oracle(x, t, ctx, added).sample
        out = eng(x, t, encoder_hidden_states=ctx, added_cond_kwargs=added).sample
    assert out.shape == ref.shape
    err = (out - ref).abs().max().item()
    assert err < 2e-4 * ref.abs().max().item() + 1e-5, err


@pytest.mark.parametrize("arch", ["tiny21", "tiny15", "tinyxl"])
def test_lora_forward_and_grads_match_oracle_autograd(arch):
    oracle = build_unet(arch)
    eng = _engine(arch, oracle)
    x, ctx, added = _inputs(arch, n=2, hw=8)
    t = torch.tensor(261)

    def make(unet):
        torch.manual_seed(11)
        with contextlib.redirect_stdout(io.StringIO()):
            net = leco_ref.LoRANetworkRef(unet, rank=4, multiplier=1.0, alpha=1.0)
        g = torch.Generator().manual_seed(5)
        for l in net.unet_loras:  # lora_up is zero at init: randomise so the branch contributes
            l.lora_up.weight.data = 0.05 * torch.randn(l.lora_up.weight.shape, generator=g)
        return net

    net_o, net_e = make(oracle), make(eng)
    # 16 Transformer2DModel x 12 adapted layers (SD1.x/2.x); tinyxl: 11 Transformer2DModel, 28 blocks: 11*2 + 28*10
    assert len(net_o.unet_loras) == len(net_e.unet_loras) == (302 if arch == "tinyxl" else 192)
    assert [l.lora_name for l in net_o.unet_loras] == [l.lora_name for l in net_e.unet_loras]
    goal = torch.randn((2, 4, 8, 8), generator=torch.Generator().manual_seed(9))

    def loss_of(unet, net):
        with net:
            y = unet(x, t, encoder_hidden_states=ctx, added_cond_kwargs=added).sample
        return torch.nn.functional.mse_loss(y.float(), goal), y

    lo, yo = loss_of(oracle, net_o)
    le, ye = loss_of(eng, net_e)
    assert (yo - ye).abs().max().item() < 2e-4 * yo.abs().max().item() + 1e-5
    lo.backward()
    le.backward()
    worst = 0.0
    for a, b in zip(net_o.unet_loras, net_e.unet_loras):
        for pa, pb in ((a.lora_down.weight, b.lora_down.weight), (a.lora_up.weight, b.lora_up.weight)):
            assert pb.grad is not None, b.lora_name
            scale = pa.grad.abs().max().item() + 1e-8
            worst = max(worst, (pa.grad - pb.grad).abs().max().item() / scale)
    assert worst < 5e-3, worst
    # multiplier == 0 (outside `with network:`) must equal the frozen network exactly
    with torch.no_grad():
        base = _engine(arch, build_unet(arch))(x, t, encoder_hidden_states=ctx, added_cond_kwargs=added).sample
        off = eng(x, t, encoder_hidden_states=ctx, added_cond_kwargs=added).sample
    assert torch.equal(base, off)


def test_c3lier_conv_adapters_match_oracle_autograd():
    """BASELINE config 3 topology (attention + conv adapters, rank 8): 278 modules incl. 3x3
    conv, stride-2 conv, 1x1 shortcut and time_emb_proj adapters — forward and all gradients vs autograd."""
    arch = "tiny15"
    oracle = build_unet(arch)
    eng = _engine(arch, oracle)
    x, ctx, _ = _inputs(arch, n=2, hw=8)
    t = torch.tensor(261)
    targets = leco_ref.ATTN_TARGETS + leco_ref.CONV_TARGETS

    def make(unet):
        torch.manual_seed(11)
        with contextlib.redirect_stdout(io.StringIO()):
            net = leco_ref.LoRANetworkRef(unet, rank=8, multiplier=1.0, alpha=4.0, targets=targets)
        g = torch.Generator().manual_seed(5)
        for l in net.unet_loras:
            l.lora_up.weight.data = 0.05 * torch.randn(l.lora_up.weight.shape, generator=g)
        return net

    net_o, net_e = make(oracle), make(eng)
    assert len(net_o.unet_loras) == len(net_e.unet_loras) == 278
    goal = torch.randn((2, 4, 8, 8), generator=torch.Generator().manual_seed(9))
    outs = []
    for unet, net in ((oracle, net_o), (eng, net_e)):
        with net:
            y = unet(x, t, encoder_hidden_states=ctx).sample
        torch.nn.functional.mse_loss(y.float(), goal).backward()
        outs.append(y.detach())
    assert (outs[0] - outs[1]).abs().max().item() < 2e-4 * outs[0].abs().max().item() + 1e-5
    worst, worst_name = 0.0, None
    for a, b in zip(net_o.unet_loras, net_e.unet_loras):
        for pa, pb in ((a.lora_down.weight, b.lora_down.weight), (a.lora_up.weight, b.lora_up.weight)):
            assert pb.grad is not None, b.lora_name
            rel = (pa.grad - pb.grad).abs().max().item() / (pa.grad.abs().max().item() + 1e-8)
            if rel > worst:
                worst, worst_name = rel, b.lora_name
    assert worst < 5e-3, (worst, worst_name)


def test_high_rank_adapters_match_oracle_autograd():
    """ranks whose stacked width exceeds one 64-wide tensor-core K-segment (rank 32: q|k|v stack 96,
    c3lier time_emb_proj groups stack 128) take the accumulate-GEMM path; forward and gradients vs autograd."""
    arch = "tiny15"
    oracle = build_unet(arch)
    eng = _engine(arch, oracle)
    x, ctx, _ = _inputs(arch, n=2, hw=8)
    t = torch.tensor(261)
    targets = leco_ref.ATTN_TARGETS + leco_ref.CONV_TARGETS

    def make(unet):
        torch.manual_seed(11)
        with contextlib.redirect_stdout(io.StringIO()):
            net = leco_ref.LoRANetworkRef(unet, rank=32, multiplier=1.0, alpha=8.0, targets=targets)
        g = torch.Generator().manual_seed(5)
        for l in net.unet_loras:
            l.lora_up.weight.data = 0.02 * torch.randn(l.lora_up.weight.shape, generator=g)
        return net

    net_o, net_e = make(oracle), make(eng)
    goal = torch.randn((2, 4, 8, 8), generator=torch.Generator().manual_seed(9))
    outs = []
    for unet, net in ((oracle, net_o), (eng, net_e)):
        with net:
            y = unet(x, t, encoder_hidden_states=ctx).sample
        torch.nn.functional.mse_loss(y.float(), goal).backward()
        outs.append(y.detach())
    assert (outs[0] - outs[1]).abs().max().item() < 2e-4 * outs[0].abs().max().item() + 1e-5
    worst, worst_name = 0.0, None
    for a, b in zip(net_o.unet_loras, net_e.unet_loras):
        for pa, pb in ((a.lora_down.weight, b.lora_down.weight), (a.lora_up.weight, b.lora_up.weight)):
            assert pb.grad is not None, b.lora_name
            rel = (pa.grad - pb.grad).abs().max().item() / (pa.grad.abs().max().item() + 1e-8)
            if rel > worst:
                worst, worst_name = rel, b.lora_name
    assert worst < 5e-3, (worst, worst_name)


@pytest.mark.parametrize("arch", ["tiny21", "tinyxl"])
def test_hoisted_cross_attention_kv_is_exact(arch):
    """`cross_kv` + `run(kv_cache=...)` (the K/V projections of the text embedding computed once per iteration) must
    reproduce the plain forward bit-for-bit, adapters on."""
    oracle = build_unet(arch)
    eng = _engine(arch, oracle)
    x, ctx, added = _inputs(arch)
    torch.manual_seed(11)
    with contextlib.redirect_stdout(io.StringIO()):
        net = leco_ref.LoRANetworkRef(eng, rank=4, multiplier=1.0, alpha=1.0)
    g = torch.Generator().manual_seed(5)
    for l in net.unet_loras:
        l.lora_up.weight.data = 0.05 * torch.randn(l.lora_up.weight.shape, generator=g)
    t = torch.full((x.shape[0],), 481.0)
    ctx2d = ctx.reshape(-1, ctx.shape[-1])
    with torch.no_grad(), net:
        eng._ensure_packed(x.device)
        plain = eng.run(x, t, ctx2d, added, None)
        kv = eng.cross_kv(ctx2d)
        assert [tuple(k.shape) for k in kv] == eng.cross_kv_shapes(ctx2d.shape[0])
        hoisted = eng.run(x, t, ctx2d, added, None, kv_cache=kv)
    assert torch.equal(plain, hoisted)


def test_repack_keeps_adapter_sites_and_sees_new_weights():
    """an in-place weight edit after a pack must invalidate the kernel-layout copies, and the re-pack must
    keep the LoraSite objects an adapter network is bound to."""
    arch = "tiny21"
    oracle = build_unet(arch)
    eng = _engine(arch, oracle)
    x, ctx, _ = _inputs(arch)
    t = torch.tensor(481)
    with torch.no_grad():
        y0 = eng(x, t, encoder_hidden_states=ctx).sample
    sites0 = eng.lora_sites()
    other = build_unet(arch, seed=1)
    eng.load_state_dict(other.state_dict())
    with torch.no_grad():
        y1 = eng(x, t, encoder_hidden_states=ctx).sample
        ref = other(x, t, ctx).sample
    assert all(a is b for a, b in zip(sites0, eng.lora_sites()))
    assert not torch.allclose(y0, y1)
    assert (y1 - ref).abs().max().item() < 2e-4 * ref.abs().max().item() + 1e-5


def test_reference_loop_body_on_the_engine_equals_the_oracle_iteration():
    """The reference's loop body (oracle/leco_ref.leco_iteration = train_lora.py:141-302, pinned against the reference)
    driving the ENGINE tree through `unet(...).sample`, `with network:`, `loss.backward()`, torch AdamW — the drop-in seam,
    here with the torch double in fp32 — gives the oracle's k draws, losses and adapter weights (GPU twin:
    tests/test_gpu_parity.py::test_reference_loop_body_drives_the_engine_through_the_drop_in_surface)."""
    from leco_b200 import lora as plora
    from oracle.sched_ref import create_noise_scheduler
    from leco_b200.synthetic import prompt_embedding
    arch = "tiny21"
    D = CONFIGS[arch].cross_attention_dim
    emb = {p: prompt_embedding(p, D) for p in ("van gogh", "")}

    def run(engine: bool):
        oracle = build_unet(arch, seed=0)
        unet = _engine(arch, oracle) if engine else oracle
        torch.manual_seed(1234)
        with contextlib.redirect_stdout(io.StringIO()):
            net = (plora.LoRANetwork if engine else leco_ref.LoRANetworkRef)(unet, rank=4, multiplier=1.0, alpha=1.0)
        pair = leco_ref.PromptPairRef(target=emb["van gogh"], positive=emb["van gogh"], unconditional=emb[""],
                                      neutral=emb[""], guidance_scale=1.0, resolution=64, batch_size=2, action="erase")
        opt = torch.optim.AdamW(net.prepare_optimizer_params(), lr=1e-3)
        lrs = torch.optim.lr_scheduler.ConstantLR(opt, factor=1)
        sched = create_noise_scheduler("ddim", "v_prediction")
        torch.manual_seed(7)
        losses, ks = [], []
        for _ in range(3):
            rec = {}
            losses.append(leco_ref.leco_iteration(unet, sched, net, opt, lrs, [pair], max_denoising_steps=6, record=rec))
            ks.append(rec["k"])
        return losses, ks, [l.lora_up.weight.detach().clone() for l in net.unet_loras]
    la, ka, ua = run(engine=True)
    lb, kb, ub = run(engine=False)
    assert ka == kb
    for a, b in zip(la, lb):
        assert abs(a - b) <= 2e-4 * abs(b), (la, lb)
    num = sum((x - y).pow(2).sum().item() for x, y in zip(ua, ub)) ** 0.5
    den = sum(y.pow(2).sum().item() for y in ub) ** 0.5
    assert den > 0 and num / den < 2e-2, num / den


def _stub_flat_kernels(monkeypatch):
    """bind_flat / FlatState call three kernels through leco_b200.ops; on the CPU they are replaced by their documented
    semantics (include/leco_b200.h: leco_transpose_tiles, leco_cast_f32_to_bf16) so the host-side bookkeeping — which
    buffer the Parameters live in, what the kernels are handed — can be checked without a GPU."""
    import struct
    from leco_b200 import ops

    def transpose_tiles(src, dst, tiles, n_tiles):
        raw = bytes(tiles.cpu().numpy().tobytes())
        seen = set()
        for i in range(n_tiles):
            so, do, rows, cols, _, _ = struct.unpack_from("<qqiiii", raw, 32 * i)
            if (so, rows, cols) in seen:
                continue
            seen.add((so, rows, cols))
            dst[do:do + rows * cols].copy_(src[so:so + rows * cols].view(rows, cols).t().reshape(-1))

    def cast_f32_to_bf16(x, out=None):
        y = x.bfloat16()
        if out is None:
            return y
        out.copy_(y)
        return out

    monkeypatch.setattr(ops, "transpose_tiles", transpose_tiles)
    monkeypatch.setattr(ops, "cast_f32_to_bf16", cast_f32_to_bf16)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32], ids=["bfloat16", "float32_master"])
def test_flat_layout_bookkeeping(monkeypatch, dtype):
    """leco_b200.lora.bind_flat on the c3lier topology (linear, 1x1 and 3x3 adapters, fused q/k/v sites): Parameters
    become views of ONE flat buffer — the bf16 operand buffer for a bf16 network, the fp32 master for a float32 network
    (`train.precision: float32`) whose bf16 operand copy is re-derived before every pass — values, forward, gradients and
    the transposed operands unchanged by the re-homing."""
    import leco_b200.lora as plora
    _stub_flat_kernels(monkeypatch)
    arch = "tiny15"
    eng = _engine(arch, build_unet(arch))
    x, ctx, _ = _inputs(arch, n=2, hw=8)
    t = torch.tensor(261)
    saved = list(plora.DEFAULT_TARGET_REPLACE)
    try:
        plora.DEFAULT_TARGET_REPLACE += plora.UNET_TARGET_REPLACE_MODULE_CONV
        torch.manual_seed(11)
        with contextlib.redirect_stdout(io.StringIO()):
            net = plora.LoRANetwork(eng, rank=4, multiplier=1.0, alpha=1.0)
    finally:
        plora.DEFAULT_TARGET_REPLACE[:] = saved
    g = torch.Generator().manual_seed(2)
    for l in net.unet_loras:
        l.lora_up.weight.data = (torch.randn(l.lora_up.weight.shape, generator=g) * 0.05)
    net.to(dtype=dtype)
    for l in net.unet_loras:        # bf16-representable values: the operand copy of the float32 network is then exact
        l.lora_down.weight.data = l.lora_down.weight.data.bfloat16().to(dtype)
        l.lora_up.weight.data = l.lora_up.weight.data.bfloat16().to(dtype)
    before = {k: v.detach().clone() for k, v in net.state_dict().items()}

    def loss_and_grads():
        for p in net.parameters():
            p.grad = None
        with net:
            y = eng(x, t, encoder_hidden_states=ctx).sample
            (y.float() ** 2).mean().backward()
        return y.detach(), [p.grad.detach().clone() for p in net.parameters()]

    y0, g0 = loss_and_grads()                 # packing path: no flat layout yet
    assert net.flat is None and all(gr.dtype == dtype for gr in g0)

    flat = net.bind_flat()
    assert (flat.master is None) == (dtype == torch.bfloat16)
    home = flat.home
    assert home.dtype == dtype and flat.params.dtype == torch.bfloat16 and flat.grads.dtype == torch.float32
    lo, hi = home.data_ptr(), home.data_ptr() + home.numel() * home.element_size()
    for k, v in net.state_dict().items():
        if k.endswith("weight"):
            assert lo <= v.data_ptr() < hi, k                       # a view of the flat buffer
        assert torch.equal(v, before[k]), k                          # same values, same shapes
    assert torch.equal(flat.params, home.bfloat16())                 # what the kernels read
    assert int(flat.mask.sum()) == sum(p.numel() for p in net.parameters()) == flat.n_real
    assert torch.equal(home[flat.mask == 0], torch.zeros_like(home[flat.mask == 0]))   # operand padding stays zero
    for s in eng.lora_sites():
        if s.adapters() is not None:
            assert s._native and s.ad.dtype == torch.bfloat16
            assert torch.equal(s.static_t[0], s.ad.t()) and torch.equal(s.static_t[1], s.bup.t())

    y1, g1 = loss_and_grads()                 # native path: operands and fp32 gradient accumulators inside the flat buffers
    assert all(s._native for s in eng.lora_sites() if s.adapters() is not None)      # still bound after a pass
    assert torch.allclose(y1, y0, rtol=1e-4, atol=1e-5)
    for a, b in zip(g1, g0):
        assert a.dtype == dtype
        assert torch.allclose(a.float(), b.float(), rtol=2e-2, atol=1e-6 + 1e-2 * b.float().abs().max().item())

    # an in-place edit by ANY optimizer (here: by hand) reaches the next pass, with or without autograd
    with torch.no_grad():
        net.unet_loras[0].lora_up.weight.add_(0.25)
        net.unet_loras[-1].lora_down.weight.mul_(2.0)
        with net:
            y2 = eng(x, t, encoder_hidden_states=ctx).sample
    assert torch.equal(flat.params, home.bfloat16())
    assert not torch.allclose(y2, y1)
    assert torch.equal(net.unet_loras[0].lora_up.weight, before[net.unet_loras[0].lora_name + ".lora_up.weight"] + 0.25)

    # the optimizer object steps the buffer the Parameters live in, with fp32 moments for a float32 network
    opt = plora.FlatOptimizer(flat, "adamw", lr=1e-3)
    assert opt.param_groups[0]["params"][0] is home
    assert opt.exp_avg.dtype == (torch.float32 if dtype == torch.float32 else torch.bfloat16)


def test_flat_layout_rejects_float16_adapters(monkeypatch):
    import leco_b200.lora as plora
    _stub_flat_kernels(monkeypatch)
    eng = _engine("tiny21", build_unet("tiny21"))
    with contextlib.redirect_stdout(io.StringIO()):
        net = plora.LoRANetwork(eng, rank=4).to(dtype=torch.float16)
    with pytest.raises(NotImplementedError, match="float16"):
        net.bind_flat()


from oracle.ref_loader import reference_available  # noqa: E402


@pytest.mark.skipif(not reference_available(), reason="reference sources only exist in the build container")
@pytest.mark.parametrize("xl", [False, True], ids=["train_lora", "train_lora_xl"])
def test_the_reference_train_py_runs_unmodified_on_the_engine(xl):
    """VERDICT r1 row ns5 with the reference's OWN files: `/root/reference/train_lora.py::train(config, prompts)` — its
    config / prompt parsing, its `LoRANetwork` (lora.py) patching the engine's tree, its `train_util.diffusion` /
    `predict_noise`, `PromptEmbedsPair.loss`, `loss.backward()`, torch AdamW, `save_weights` — runs UNMODIFIED with
    `model_util.load_models` returning the engine (the one binding of INTEGRATION.md section 1; here the torch double in
    fp32, `train.precision: float32`).  Same k draws, losses and saved adapter file as the same function on the oracle
    UNet (= the run tests/golden/leco_train_golden.json was made from).  train_lora_xl: `train_lora_xl.py::train` the
    same way (pooled text embedding + add_time_ids through `added_cond_kwargs`; the harness works around the two
    reference bugs tests/golden/make_golden.py documents)."""
    from tests.golden import make_golden as mg
    arch, iters, max_steps = ("tinyxl", 3, 6) if xl else ("tiny21", 3, 6)
    run = (lambda **kw: mg.run_reference_train_xl(arch, iters, max_steps, **kw)) if xl else \
        (lambda **kw: mg.run_reference_train(arch, iters, max_steps, True, **kw))
    ref_losses, ref_ks, ref_sd, _ = run()
    losses, ks, sd, _ = run(unet=_engine(arch, build_unet(arch, seed=0)))
    assert ks == ref_ks and len(losses) == len(ref_losses) == iters
    for a, b in zip(losses, ref_losses):
        assert abs(a - b) <= 2e-4 * abs(b), (losses, ref_losses)
    assert list(sd) == list(ref_sd) and len(sd) % 3 == 0 and len(sd) >= 3 * 16     # kohya key set in the reference's order
    num = sum((sd[k].float() - ref_sd[k].float()).pow(2).sum().item() for k in sd if k.endswith("lora_up.weight")) ** 0.5
    den = sum(ref_sd[k].float().pow(2).sum().item() for k in sd if k.endswith("lora_up.weight")) ** 0.5
    assert den > 0 and num / den < 2e-2, num / den
    for k in sd:
        if k.endswith("lora_down.weight"):
            assert torch.allclose(sd[k], ref_sd[k], rtol=1e-3, atol=1e-4), k
