"""GPU checks of every non-GEMM kernel and of the engine end to end (run through gpurun).

Each case runs in a subprocess under a timeout.  References: tests/torch_backend.py (plain
PyTorch fp32 of the same op) and the oracle (fp32 CPU UNet + autograd).
Results -> gpurun_out/kernel_cases.json
"""
from __future__ import annotations

import contextlib
import io
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def _rand(shape, scale=1.0, seed=0, dtype=None):
    import torch
    g = torch.Generator(device="cpu").manual_seed(seed)
    x = torch.randn(shape, generator=g) * scale
    return x.to(dtype or torch.bfloat16).cuda()


def _cmp(out, ref, tol=2e-2):
    import torch
    out, ref = out.float().cpu(), ref.float().cpu()
    err = (out - ref).abs().max().item()
    scale = ref.abs().max().item() + 1e-6
    return {"max_abs_err": err, "ref_absmax": scale, "rel": err / scale,
            "ok": bool(torch.isfinite(out).all().item()) and err / scale < tol}


def _merge(results):
    ok = all(r["ok"] for r in results.values())
    return {"ok": ok, "rel": max(r["rel"] for r in results.values()), "parts": results}


def case_norms():
    import torch
    from leco_b200 import ops
    from tests import torch_backend as tb
    res = {}
    # (2, 100, 320): rows not a multiple of the cluster size (ragged / empty CTAs of the one-launch cluster kernel);
    # (1, 4096, 640): a sample too large for one cluster's shared memory (two-launch path inside leco_group_norm_v3)
    for (n, hw, c) in ((2, 4096, 320), (3, 256, 1280), (2, 64, 2560), (4, 4, 64), (2, 1024, 960), (2, 100, 320),
                       (1, 4096, 640), (5, 1024, 1280)):
        x = _rand((n * hw, c), seed=1) * 1.5 + 0.3
        gm, bt = _rand((c,), 0.2, 2) + 1.0, _rand((c,), 0.2, 3)
        for silu in (True, False):
            y, st = ops.group_norm(x, n, hw, gm, bt, 32, 1e-5, silu)
            yr, str_ = tb.group_norm(x.cpu(), n, hw, gm.cpu(), bt.cpu(), 32, 1e-5, silu)
            res[f"gn_{n}_{hw}_{c}_{int(silu)}"] = _cmp(y, yr)
            res[f"gn_stats_{n}_{hw}_{c}_{int(silu)}"] = _cmp(st, str_, 1e-2)
            dz = _rand((n * hw, c), seed=4)
            dx = ops.group_norm_bwd(x, dz, st, gm, bt, n, hw, 32, silu)
            dxr = tb.group_norm_bwd(x.cpu(), dz.cpu(), str_, gm.cpu(), bt.cpu(), n, hw, 32, silu)
            res[f"gn_bwd_{n}_{hw}_{c}_{int(silu)}"] = _cmp(dx, dxr, 3e-2)
    for (m, c) in ((16384, 320), (1024, 1280), (77, 64), (300, 640)):
        x = _rand((m, c), seed=5) * 2 - 0.5
        gm, bt = _rand((c,), 0.2, 6) + 1.0, _rand((c,), 0.2, 7)
        y, st = ops.layer_norm(x, gm, bt, 1e-5, True)
        yr, str_ = tb.layer_norm(x.cpu(), gm.cpu(), bt.cpu(), 1e-5, True)
        res[f"ln_{m}_{c}"] = _cmp(y, yr)
        dy = _rand((m, c), seed=8)
        dx = ops.layer_norm_bwd(x, dy, st, gm)
        res[f"ln_bwd_{m}_{c}"] = _cmp(dx, tb.layer_norm_bwd(x.cpu(), dy.cpu(), str_, gm.cpu()), 3e-2)
    return _merge(res)


def case_norm_perf():
    """Perf triage (not a parity case): GroupNorm / LayerNorm at the SD2.1 denoise-forward shapes (4 samples), timed in a
    CUDA graph of 20 back-to-back launches (warm L2, like the real forward), vs the HBM roofline of 3 passes x 2 B."""
    import torch
    from leco_b200 import ops
    res = {"ok": True}
    for (n, hw, c, silu) in ((4, 4096, 320, True), (4, 4096, 640, True), (4, 4096, 960, True), (4, 1024, 640, True),
                             (4, 1024, 1280, True), (4, 1024, 1920, True), (4, 256, 1280, True), (4, 256, 2560, True),
                             (4, 64, 1280, True)):
        x = _rand((n * hw, c), seed=1)
        gm, bt = _rand((c,), 0.2, 2) + 1.0, _rand((c,), 0.2, 3)
        for _ in range(3):
            ops.group_norm(x, n, hw, gm, bt, 32, 1e-5, silu)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(20):
                ops.group_norm(x, n, hw, gm, bt, 32, 1e-5, silu)
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 100 * 1e3
        res[f"gn_{hw}_{c}_us"] = round(us, 2)
        res[f"gn_{hw}_{c}_gbs"] = round(3 * 2 * n * hw * c / us / 1e3, 1)
    for (m, c) in ((16384, 320), (4096, 640), (1024, 1280)):
        x = _rand((m, c), seed=5)
        gm, bt = _rand((c,), 0.2, 6) + 1.0, _rand((c,), 0.2, 7)
        for _ in range(3):
            ops.layer_norm(x, gm, bt, 1e-5, False)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(20):
                ops.layer_norm(x, gm, bt, 1e-5, False)
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 100 * 1e3
        res[f"ln_{m}_{c}_us"] = round(us, 2)
        res[f"ln_{m}_{c}_gbs"] = round(2 * 2 * m * c / us / 1e3, 1)
    return res


def case_elementwise():
    import torch
    from leco_b200 import ops
    from tests import torch_backend as tb
    tb.ACT_DTYPE = torch.bfloat16
    res = {}
    x = _rand((3, 4, 16, 24), seed=1, dtype=torch.float32)
    w, b = _rand((64, 4, 3, 3), 0.3, 2), _rand((64,), 0.1, 3)
    res["conv_in_f32"] = _cmp(ops.conv_in(x, w, b), tb.conv_in(x.cpu(), w.cpu(), b.cpu()))
    res["conv_in_bf16"] = _cmp(ops.conv_in(x.bfloat16(), w, b), tb.conv_in(x.bfloat16().cpu(), w.cpu(), b.cpu()))
    # width not a multiple of 4: the one-pixel-per-thread kernel; 64x64 -> 320: the UNet's own conv_in shape
    x2 = _rand((2, 4, 10, 18), seed=8, dtype=torch.float32)
    res["conv_in_w18"] = _cmp(ops.conv_in(x2, w, b), tb.conv_in(x2.cpu(), w.cpu(), b.cpu()))
    x3 = _rand((2, 4, 64, 64), seed=9, dtype=torch.float32)
    w3, b3 = _rand((320, 4, 3, 3), 0.3, 10), _rand((320,), 0.1, 11)
    res["conv_in_64_320"] = _cmp(ops.conv_in(x3, w3, b3), tb.conv_in(x3.cpu(), w3.cpu(), b3.cpu()))
    xa = _rand((3 * 16 * 24, 64), seed=4)
    wo, bo = _rand((4, 9, 64), 0.05, 5), _rand((4,), 0.1, 6)
    res["conv_out"] = _cmp(ops.conv_out(xa, wo, bo, 3, 16, 24), tb.conv_out(xa.cpu(), wo.cpu(), bo.cpu(), 3, 16, 24))
    dy = _rand((3, 4, 16, 24), seed=7, dtype=torch.float32)
    res["conv_out_bwd"] = _cmp(ops.conv_out_bwd(dy, wo, 64), tb.conv_out_bwd(dy.cpu(), wo.cpu(), 64))
    t = torch.tensor([999.0, 481.0, 1.0, 20.0], device="cuda")
    res["temb"] = _cmp(ops.timestep_embedding(t, 320), tb.timestep_embedding(t.cpu(), 320))
    v = _rand((1000, 64), seed=8)
    res["silu"] = _cmp(ops.silu(v), tb.silu(v.cpu()))
    y = v.clone()
    res["add"] = _cmp(ops.add_(y, v), tb.add_(v.cpu().clone(), v.cpu()))
    pre = _rand((300, 512), seed=9)
    res["geglu"] = _cmp(ops.geglu_fwd(pre), tb.geglu_fwd(pre.cpu()))
    do = _rand((300, 256), seed=10)
    res["geglu_bwd"] = _cmp(ops.geglu_bwd(pre, do), tb.geglu_bwd(pre.cpu(), do.cpu()))
    a, b2 = _rand((500, 64), seed=11), _rand((500, 192), seed=12)
    cat = ops.concat2(a, b2)
    res["concat"] = _cmp(cat, tb.concat2(a.cpu(), b2.cpu()), 1e-6)
    sa, sb = ops.split2(cat, 64)
    res["split_a"] = _cmp(sa, a, 1e-6)
    res["split_b"] = _cmp(sb, b2, 1e-6)
    xi = _rand((2 * 8 * 12, 64), seed=13)
    res["upsample"] = _cmp(ops.upsample2x(xi, 2, 8, 12), tb.upsample2x(xi.cpu(), 2, 8, 12), 1e-6)
    dyu = _rand((2 * 16 * 24, 64), seed=14)
    res["upsample_bwd"] = _cmp(ops.upsample2x_bwd(dyu, 2, 8, 12), tb.upsample2x_bwd(dyu.cpu(), 2, 8, 12))
    xs = _rand((2 * 8 * 12, 64), seed=15)
    res["im2col_s2"] = _cmp(ops.im2col_s2(xs, 2, 8, 12), tb.im2col_s2(xs.cpu(), 2, 8, 12), 1e-6)
    dc = _rand((2 * 4 * 6, 9 * 64), seed=16)
    res["col2im_s2"] = _cmp(ops.col2im_s2(dc, 2, 8, 12), tb.col2im_s2(dc.cpu(), 2, 8, 12))
    src = _rand((2, 3, 77, 64), seed=17)
    res["transpose_pad"] = _cmp(ops.transpose_batched(src, cols_pad=80), tb.transpose_batched(src.cpu(), 80), 1e-6)
    wide = _rand((2 * 100, 3 * 128), seed=18)
    view = wide[:, 128:256].unflatten(0, (2, 100)).unflatten(2, (2, 64)).permute(0, 2, 1, 3)
    res["transpose_strided"] = _cmp(ops.transpose_batched(view), tb.transpose_batched(view.cpu()), 1e-6)
    s = _rand((2, 2, 50, 80), 3.0, 19, torch.float32)
    res["softmax"] = _cmp(ops.softmax_rows(s, 77, 80),
                          torch.nn.functional.pad(torch.softmax(s.cpu()[..., :77], -1), (0, 3)))
    p = ops.softmax_rows(s, 77, 80)
    dp = _rand((2, 2, 50, 80), 1.0, 20, torch.float32)
    pf = p.float().cpu()[..., :77]
    dref = pf * (dp.cpu()[..., :77] - (pf * dp.cpu()[..., :77]).sum(-1, keepdim=True)) * 0.125
    res["softmax_bwd"] = _cmp(ops.softmax_bwd_rows(p, dp, 77, 0.125), torch.nn.functional.pad(dref, (0, 3)))
    return _merge(res)


def case_training_kernels():
    import torch
    from leco_b200 import ops
    res = {}
    for (M, N1, N2, tr) in ((16384, 960, 16, False), (1000, 320, 16, True), (4096, 2560, 32, False), (308, 1024, 16, True),
                            (16384, 320, 64, False), (77, 1280, 48, True), (33, 200, 8, False), (2048, 328, 24, True)):
        a, b = _rand((M, N1), seed=1), _rand((M, N2), seed=2)
        out = torch.full((N2, N1) if tr else (N1, N2), 0.25, device="cuda")      # the kernel accumulates into `out`
        ops.tn_reduce(a, b, out, 0.5, transpose_out=tr)
        ref = 0.25 + 0.5 * a.float().t() @ b.float()
        res[f"tn_{M}_{N1}_{N2}_{int(tr)}"] = _cmp(out, ref.t() if tr else ref, 2e-3)
    # strided operands (column slices of wider buffers, as the backward passes them)
    wide_a, wide_b = _rand((4096, 3 * 320), seed=5), _rand((4096, 96), seed=6)
    out = torch.zeros((320, 32), device="cuda")
    ops.tn_reduce(wide_a[:, 320:640], wide_b[:, 32:64], out, 1.0)
    res["tn_strided"] = _cmp(out, wide_a[:, 320:640].float().t() @ wide_b[:, 32:64].float(), 2e-3)
    t0 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a, b = _rand((16384, 320), seed=1), _rand((16384, 16), seed=2)
    out = torch.zeros((320, 16), device="cuda")
    for _ in range(3):
        ops.tn_reduce(a, b, out, 1.0)
    t0[0].record()
    for _ in range(20):
        ops.tn_reduce(a, b, out, 1.0)
    t0[1].record()
    torch.cuda.synchronize()
    res["tn_strided"]["tn_16384x320x16_us"] = t0[0].elapsed_time(t0[1]) * 50.0
    # fused AdamW vs torch.optim.AdamW on fp32 (bf16 rounding of params bounds the difference)
    n = 100000
    p0 = _rand((n,), 0.1, 3)
    g = _rand((n,), 0.01, 4, torch.float32)
    p = p0.clone()
    m = torch.zeros(n, device="cuda")
    v = torch.zeros(n, device="cuda")
    hyper = torch.tensor([1e-3, 0.9, 0.999, 1e-8, 0.01, 1.0, 1.0, 0.0], device="cuda")
    pt = torch.nn.Parameter(p0.float().clone())
    opt = torch.optim.AdamW([pt], lr=1e-3)
    for step in range(1, 4):
        hyper[5] = float(step)
        gg = g.clone() * step
        ops.adamw_flat(p, gg, m, v, None, hyper, zero_grad=True)
        pt.grad = (g * step).bfloat16().float()
        opt.step()
    res["adamw_3steps"] = _cmp(p, pt.data, 1e-2)
    # parameters are STORED in bf16 (like the reference's): the result may differ from fp32 AdamW by
    # about one bf16 ulp of the parameter, and must move in the same direction
    ulp = pt.data.abs().clamp_min(1e-3) * 2.0 ** -7
    worst = ((p.float() - pt.data).abs() / ulp).max().item()
    res["adamw_within_ulps"] = {"rel": worst, "ok": worst < 3.0, "max_abs_err": worst, "ref_absmax": 1.0}
    cos = torch.nn.functional.cosine_similarity((p.float() - p0.float()).flatten(), (pt.data - p0.float()).flatten(), dim=0).item()
    res["adamw_update_cos"] = {"rel": 1 - cos, "ok": cos > 0.9, "max_abs_err": 1 - cos, "ref_absmax": 1.0}
    # guided step + loss
    eps = _rand((4, 4, 8, 8), seed=5, dtype=torch.float32)
    x = _rand((2, 4, 8, 8), seed=6, dtype=torch.float32)
    coef = torch.tensor([3.0, 1.01, -0.07], device="cuda")
    xo, gd = ops.guided_step(eps, x, coef, True, True)
    gref = eps[:2] + 3.0 * (eps[2:] - eps[:2])
    res["guided"] = _cmp(gd, gref, 1e-5)
    res["ddim"] = _cmp(xo, 1.01 * x - 0.07 * gref, 1e-5)
    t_, p_, n_, u_ = (_rand((2, 4, 8, 8), seed=s, dtype=torch.float32) for s in (7, 8, 9, 10))
    loss, dt = ops.leco_loss(t_, p_, n_, u_, -1.5)
    goal = n_ - 1.5 * (p_ - u_)
    res["loss"] = _cmp(loss, ((t_ - goal) ** 2).mean().reshape(1), 1e-5)
    res["dloss"] = _cmp(dt, 2 * (t_ - goal) / t_.numel(), 1e-5)
    return _merge(res)


def case_sched_step():
    """leco_sched_step / leco_scale_by_dev vs the coefficient-row formula (pinned against the oracle's DDPM / LMS /
    Euler-a restatements on the CPU by tests/test_host_logic_cpu.py) over a 6-step chain incl. the LMS history ring, and
    the drop-in `scheduler.step()` surface on device tensors."""
    import torch
    from leco_b200 import ops
    from leco_b200.scheduler import create_noise_scheduler
    res = {}
    g = torch.Generator().manual_seed(3)
    for name in ("ddpm", "lms", "euler_a"):
        for ptype in ("epsilon", "v_prediction"):
            s = create_noise_scheduler(name, ptype)
            s.set_timesteps(50)
            rows = torch.tensor(s.table(3.0), dtype=torch.float32, device="cuda")
            x = (torch.randn((2, 4, 8, 8), generator=g) * float(s.init_noise_sigma)).cuda()
            xr = x.double().cpu()
            hist = torch.zeros((4, x.numel()), device="cuda") if s.history else None
            hist_r = []
            x_drop = x.clone()
            worst = worst_drop = 0.0
            for i in range(6):
                eps = torch.randn((4, 4, 8, 8), generator=g).cuda()
                noise = torch.randn((2, 4, 8, 8), generator=g).cuda() if s.needs_noise else None
                row = [float(v) for v in rows[i].double().cpu()]
                gd = (eps[:2] + 3.0 * (eps[2:] - eps[:2])).double().cpu()
                xin = ops.scale_by_dev(x, rows[i], 4)
                worst = max(worst, (xin.double().cpu() - xr * row[4]).abs().max().item())
                ref = row[1] * xr + row[2] * gd + (row[3] * noise.double().cpu() if noise is not None else 0)
                if s.history:
                    hist_r.append(row[5] * xr + row[6] * gd)
                    hist_r = hist_r[-4:]
                    for c, h in zip(row[7:11], reversed(hist_r)):
                        ref = ref + c * h
                # drop-in surface: scheduler.step(model_output, t, sample) with the same guided prediction and noise
                x_drop = s.step(gd.float().cuda(), s.timesteps[i], x_drop, noise=noise).prev_sample
                ops.sched_step(eps, x, rows[i], noise=noise, hist=hist, out=x)
                scale = ref.abs().max().item() + 1e-9
                worst = max(worst, (x.double().cpu() - ref).abs().max().item() / scale)
                worst_drop = max(worst_drop, (x_drop.double().cpu() - ref).abs().max().item() / scale)
                xr = ref
            res[f"{name}_{ptype}"] = {"rel": worst, "ok": worst < 1e-5, "max_abs_err": worst, "ref_absmax": 1.0}
            res[f"{name}_{ptype}_dropin"] = {"rel": worst_drop, "ok": worst_drop < 1e-5, "max_abs_err": worst_drop, "ref_absmax": 1.0}
    return _merge(res)


def _bf16_ulp_stats(got, ref, floor=1e-3):
    """bf16 tensors: fraction of elements that are not bit-equal and the worst distance in bf16 ulps, an ulp being that
    of max(|ref|, floor) (parameters that pass through zero have arbitrarily small ulps of their own; `floor` = a tenth
    of a typical update keeps the measure meaningful there)."""
    import torch
    g, r = got.float(), ref.float()
    mism = (got.view(torch.int16) != ref.view(torch.int16)).float().mean().item()
    mag = torch.maximum(r.abs(), torch.full_like(r, floor))
    ulp = torch.exp2(torch.floor(torch.log2(mag)) - 7)      # bf16: 8 significant bits
    return mism, float(((g - r).abs() / ulp).max().item())


def case_optimizers():
    """leco_optim_flat with bf16 state (the benchmarked variant) vs the reference's optimizer objects on bf16
    parameters: torch.optim.AdamW / Adam (foreach) and the published lion_pytorch update rule restated with torch
    bf16 ops.  The kernel rounds where torch rounds, so the parameters must agree to the bit (a handful of 1-ulp
    differences from fused multiply-adds inside torch's kernels are tolerated)."""
    import torch
    from leco_b200 import lora as plora
    res = {}
    n = 200000
    g32 = [_rand((n,), 0.01 * (1 + s), 40 + s, torch.float32) for s in range(6)]
    mask = (torch.rand(n, generator=torch.Generator().manual_seed(3)) < 0.9).to(torch.uint8).cuda()

    def ours(name, steps, gscale=1.0, **kw):
        p = _rand((n,), 0.05, 3)
        flat = plora.FlatState(p, torch.zeros(n + 8, device="cuda"), mask)
        opt = plora.FlatOptimizer(flat, name, **kw)
        for s in range(steps):
            flat.grads.copy_(g32[s] / gscale)
            opt.step(grad_scale=gscale)
        assert float(flat.grads.abs().max()) == 0.0      # zero_grad
        return p, opt

    def theirs(make, steps):
        p0 = _rand((n,), 0.05, 3)
        pt = torch.nn.Parameter(p0.clone())
        opt = make([pt])
        for s in range(steps):
            pt.grad = (g32[s] * mask).bfloat16()
            opt.step()
        # masked-out elements (operand padding) are never touched by the fused kernel
        return torch.where(mask.bool(), pt.data, p0)

    def lion_ref(steps, lr=1e-4, b1=0.9, b2=0.99, wd=0.0):
        p = _rand((n,), 0.05, 3)
        p0 = p.clone()
        m = torch.zeros_like(p)
        for s in range(steps):
            g = g32[s].bfloat16()
            p.mul_(1 - lr * wd)
            upd = m.clone().mul_(b1).add(g, alpha=1 - b1).sign_()
            p.add_(upd, alpha=-lr)
            m.mul_(b2).add_(g, alpha=1 - b2)
        return torch.where(mask.bool(), p, p0)

    for tag, name, kw, ref in (
            ("adamw", "adamw", dict(lr=1e-3), lambda: theirs(lambda ps: torch.optim.AdamW(ps, lr=1e-3, foreach=True), 6)),
            ("adamw_lr1e-4", "adamw", dict(lr=1e-4, weight_decay=0.1),
             lambda: theirs(lambda ps: torch.optim.AdamW(ps, lr=1e-4, weight_decay=0.1, foreach=True), 6)),
            ("adam", "adam", dict(lr=1e-3), lambda: theirs(lambda ps: torch.optim.Adam(ps, lr=1e-3, foreach=True), 6)),
            ("adam_wd", "adam", dict(lr=1e-3, weight_decay=0.01),
             lambda: theirs(lambda ps: torch.optim.Adam(ps, lr=1e-3, weight_decay=0.01, foreach=True), 6)),
            ("lion", "lion", dict(lr=1e-4), lambda: lion_ref(6)),
            ("lion_wd", "lion", dict(lr=1e-4, weight_decay=0.1), lambda: lion_ref(6, wd=0.1))):
        p, _ = ours(name, 6, **kw)
        frac, ulps = _bf16_ulp_stats(p, ref())
        res[f"{tag}_bf16_state"] = {"rel": frac, "max_ulps": ulps, "ok": frac < 2e-3 and ulps <= 2, "max_abs_err": ulps,
                                    "ref_absmax": 1.0}
    # grad_scale (data-parallel mean) is applied before the bf16 rounding of the gradient
    p_a, _ = ours("adamw", 4, gscale=0.5, lr=1e-3)
    p_b, _ = ours("adamw", 4, gscale=1.0, lr=1e-3)
    frac, ulps = _bf16_ulp_stats(p_a, p_b)
    res["grad_scale"] = {"rel": frac, "max_ulps": ulps, "ok": frac < 2e-3 and ulps <= 1, "max_abs_err": ulps, "ref_absmax": 1.0}
    return _merge(res)


def case_optimizers_master():
    """leco_optim_flat_master (`train.precision: float32`: fp32 master adapters + fp32 moments) vs the reference's
    optimizer objects on fp32 parameters — torch.optim.AdamW / Adam and the lion_pytorch rule — to fp32 rounding, and
    the bf16 operand copy it writes in the same pass == the master rounded to bf16, exactly."""
    import torch
    from leco_b200 import lora as plora
    res = {}
    n = 200000
    g32 = [_rand((n,), 0.01 * (1 + s), 40 + s, torch.float32) for s in range(6)]
    mask = (torch.rand(n, generator=torch.Generator().manual_seed(3)) < 0.9).to(torch.uint8).cuda()
    p_init = _rand((n,), 0.05, 3, torch.float32)

    def ours(name, steps, gscale=1.0, **kw):
        master = p_init.clone()
        flat = plora.FlatState(master.bfloat16(), torch.zeros(n + 8, device="cuda"), mask, master=master)
        opt = plora.FlatOptimizer(flat, name, **kw)
        assert opt.exp_avg.dtype == torch.float32 and opt.param_groups[0]["params"][0] is master
        for s in range(steps):
            flat.grads.copy_(g32[s] / gscale)
            opt.step(grad_scale=gscale)
        assert float(flat.grads.abs().max()) == 0.0      # zero_grad
        return flat

    def theirs(make, steps):
        pt = torch.nn.Parameter(p_init.clone())
        opt = make([pt])
        for s in range(steps):
            pt.grad = g32[s] * mask
            opt.step()
        return torch.where(mask.bool(), pt.data, p_init)

    def lion_ref(steps, lr=1e-4, b1=0.9, b2=0.99, wd=0.0):
        p = p_init.clone()
        m = torch.zeros_like(p)
        for s in range(steps):
            g = g32[s]
            p.mul_(1 - lr * wd)
            upd = m.clone().mul_(b1).add(g, alpha=1 - b1).sign_()
            p.add_(upd, alpha=-lr)
            m.mul_(b2).add_(g, alpha=1 - b2)
        return torch.where(mask.bool(), p, p_init)

    for tag, name, kw, ref in (
            ("adamw", "adamw", dict(lr=1e-3), lambda: theirs(lambda ps: torch.optim.AdamW(ps, lr=1e-3), 6)),
            ("adamw_wd", "adamw", dict(lr=1e-4, weight_decay=0.1),
             lambda: theirs(lambda ps: torch.optim.AdamW(ps, lr=1e-4, weight_decay=0.1), 6)),
            ("adam_wd", "adam", dict(lr=1e-3, weight_decay=0.01),
             lambda: theirs(lambda ps: torch.optim.Adam(ps, lr=1e-3, weight_decay=0.01), 6)),
            ("lion_wd", "lion", dict(lr=1e-4, weight_decay=0.1), lambda: lion_ref(6, wd=0.1))):
        flat = ours(name, 6, **kw)
        want = ref()
        # the update itself (lr-sized), not the 0.05-sized weights, is what must agree: error relative to the total movement.
        # lion's sign() may flip where |u| sits at fp32 rounding level: a 1e-3 fraction of full-step disagreements is allowed
        moved = (want - p_init).abs().max().item()
        err = (flat.master - want).abs()
        rel = (err.max().item() / moved) if name != "lion" else float((err > 0.1 * moved).float().mean().item())
        shadow_exact = bool(torch.equal(flat.params, flat.master.bfloat16()))
        untouched = bool(torch.equal(flat.master[~mask.bool()], p_init[~mask.bool()]))
        res[f"{tag}_fp32_master"] = {"rel": rel, "ok": rel < 1e-3 and shadow_exact and untouched, "shadow_exact": shadow_exact,
                                     "masked_untouched": untouched, "max_abs_err": err.max().item(), "ref_absmax": moved}
    a = ours("adamw", 4, gscale=0.5, lr=1e-3)
    b = ours("adamw", 4, gscale=1.0, lr=1e-3)
    d = (a.master - b.master).abs().max().item()
    res["grad_scale"] = {"rel": d, "ok": d < 1e-7, "max_abs_err": d, "ref_absmax": 1.0}
    return _merge(res)


def case_transpose_tiles():
    """FlatState.refresh_transposed: every site's static (ad^T, bup^T) equals the transposed operand, exactly."""
    import torch
    from leco_b200.lora import LoRANetwork
    _, eng = _engine_pair("tiny15")
    torch.manual_seed(11)
    import leco_b200.lora as plora
    saved = list(plora.DEFAULT_TARGET_REPLACE)
    try:
        plora.DEFAULT_TARGET_REPLACE += plora.UNET_TARGET_REPLACE_MODULE_CONV   # c3lier: conv sites too
        with contextlib.redirect_stdout(io.StringIO()):
            net = LoRANetwork(eng, rank=8, multiplier=1.0, alpha=4.0)
    finally:
        plora.DEFAULT_TARGET_REPLACE[:] = saved
    net.to("cuda", dtype=torch.bfloat16)
    g = torch.Generator().manual_seed(5)
    with torch.no_grad():
        for l in net.unet_loras:
            l.lora_up.weight.copy_((0.05 * torch.randn(l.lora_up.weight.shape, generator=g)).to("cuda", torch.bfloat16))
    net.flat.refresh_transposed()
    torch.cuda.synchronize()
    bad = 0
    sites = [s for s in eng.lora_sites() if s.adapters() is not None]
    for s in sites:
        adT, bupT = s.static_t
        bad += int(not torch.equal(adT, s.ad.t())) + int(not torch.equal(bupT, s.bup.t()))
    return {"ok": bad == 0 and len(sites) > 100, "rel": float(bad), "sites": len(sites)}


def case_determinism():
    """leco_set_deterministic(1): two fresh trainers on the same seeds give bit-identical losses and parameters."""
    import torch
    from leco_b200 import ops
    from __graft_entry__ import engine_trainer
    ops.set_deterministic(True)
    try:
        outs = []
        for _ in range(2):
            trainer, net = engine_trainer(use_graphs=True)
            torch.manual_seed(7)
            losses = [trainer.iteration().item() for _ in range(3)]
            outs.append((losses, net.flat.params.clone()))
        same_loss = outs[0][0] == outs[1][0]
        same_p = torch.equal(outs[0][1], outs[1][1])
    finally:
        ops.set_deterministic(False)
    return {"ok": bool(same_loss and same_p), "rel": 0.0 if same_loss and same_p else 1.0, "losses": outs[0][0],
            "losses_2": outs[1][0], "params_equal": bool(same_p)}


def case_attention(nb, sq, skv, heads, d, cross=False, bwd_impl="flash"):
    """grad-pass attention (forward that saves for the backward + backward) vs plain fp32 torch.  bwd_impl "flash" =
    fused tcgen05 forward (+lse) and fused backward kernel (head dims <= 64); "v0" = the materialised path (used for
    larger head dims and under leco_set_deterministic)."""
    import torch
    from leco_b200 import ops
    from tests import torch_backend as tb
    ops.ATTENTION_BWD_IMPL = bwd_impl
    C = heads * d
    if cross:
        qbuf, kvbuf = _rand((nb * sq, C), seed=1), _rand((nb * skv, 2 * C), seed=2)
        qt, kt, vt = qbuf, kvbuf[:, :C], kvbuf[:, C:]
    else:
        qkv = _rand((nb * sq, 3 * C), seed=1)
        qt, kt, vt = qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:]
    scale = d ** -0.5
    o, saved = ops.attention(qt, kt, vt, nb, sq, skv, heads, d, scale, True)
    used = "flash" if saved[0] == "flash" else "v0"
    if bwd_impl == "flash" and d <= 64:
        assert used == "flash", "the fused backward path was not taken"
    oref, _ = tb.attention(qt.cpu(), kt.cpu(), vt.cpu(), nb, sq, skv, heads, d, scale, False)
    res = {"fwd": _cmp(o, oref)}
    go = _rand((nb * sq, C), seed=3)
    dq, dk, dv = (torch.empty((nb * s, C), device="cuda", dtype=torch.bfloat16) for s in (sq, skv, skv))
    ops.attention_bwd(go, qt, kt, vt, saved, nb, sq, skv, heads, d, scale, dq, dk, dv)
    rq, rk, rv = (torch.empty((nb * s, C)) for s in (sq, skv, skv))
    tb.attention_bwd(go.cpu(), qt.cpu(), kt.cpu(), vt.cpu(), None, nb, sq, skv, heads, d, scale, rq, rk, rv)
    res["dq"], res["dk"], res["dv"] = _cmp(dq, rq, 3e-2), _cmp(dk, rk, 3e-2), _cmp(dv, rv, 3e-2)
    return _merge(res)


def case_flash(nb, sq, skv, heads, d, cross=False, v_mode=0, perf=False, ts=None, qscale=1.0, kramp=0.0):
    """ts: LECO_FLASH_TS for this process (None = library default, 0 = the shared-memory-P variant).  qscale / kramp
    stretch the scores (and make later keys larger) so that the running max keeps moving by more than the lazy-rescale
    threshold: the in-TMEM accumulator rescale of the default variant is then taken on most tiles."""
    import torch
    if ts is not None:
        os.environ["LECO_FLASH_TS"] = str(ts)
    from leco_b200 import ops
    from tests import torch_backend as tb
    C = heads * d
    if cross:
        qbuf, kvbuf = _rand((nb * sq, C), seed=1), _rand((nb * skv, 2 * C), seed=2)
        qt, kt, vt = qbuf, kvbuf[:, :C], kvbuf[:, C:]
    else:
        qkv = _rand((nb * sq, 3 * C), seed=1)
        if qscale != 1.0 or kramp:
            q32 = qkv.float()
            q32[:, :C] *= qscale
            if kramp:
                ramp = 0.5 + kramp * (torch.arange(nb * sq, device=q32.device) % skv).float() / skv
                q32[:, C:2 * C] *= ramp[:, None]
            qkv = q32.to(torch.bfloat16)
        qt, kt, vt = qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:]
    scale = d ** -0.5
    o = ops.flash_attention(qt, kt, vt, nb, sq, skv, heads, d, scale, v_mode=v_mode)
    torch.cuda.synchronize()
    if nb * heads * sq * skv <= 2 * 8 * 1024 * 1024 * 4:
        oref, _ = tb.attention(qt.cpu(), kt.cpu(), vt.cpu(), nb, sq, skv, heads, d, scale, False)
    else:
        # big case: an INDEPENDENT reference — stock torch attention in fp32 on the GPU (the fused
        # bf16 backends do not take fp32, so this is torch's own math / memory-efficient path), one sample at a time
        def heads_view(t, s):
            return t.unflatten(0, (nb, s)).unflatten(2, (heads, d)).permute(0, 2, 1, 3).float()
        q4, k4, v4 = heads_view(qt, sq), heads_view(kt, skv), heads_view(vt, skv)
        oref = torch.cat([torch.nn.functional.scaled_dot_product_attention(q4[b:b + 1], k4[b:b + 1], v4[b:b + 1], scale=scale)
                          for b in range(nb)], 0).permute(0, 2, 1, 3).reshape(nb * sq, C)
        del q4, k4, v4
    res = _cmp(o, oref)
    if perf:
        def timeit(fn, iters=10):
            for _ in range(2):
                fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(iters):
                fn()
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / iters
        flops = 4.0 * nb * heads * sq * skv * d
        ms = timeit(lambda: ops.flash_attention(qt, kt, vt, nb, sq, skv, heads, d, scale, v_mode=v_mode))
        ms0 = timeit(lambda: ops.attention_v0(qt, kt, vt, nb, sq, skv, heads, d, scale, False), 3)
        q4 = qt.unflatten(0, (nb, sq)).unflatten(2, (heads, d)).permute(0, 2, 1, 3)
        k4 = kt.unflatten(0, (nb, skv)).unflatten(2, (heads, d)).permute(0, 2, 1, 3)
        v4 = vt.unflatten(0, (nb, skv)).unflatten(2, (heads, d)).permute(0, 2, 1, 3)
        ms_sdpa = timeit(lambda: torch.nn.functional.scaled_dot_product_attention(q4, k4, v4))
        res.update(flash_ms=ms, flash_tflops=flops / ms / 1e9, v0_ms=ms0, torch_sdpa_ms=ms_sdpa)
    return res


def _engine_pair(arch, dtype=None):
    import torch
    from leco_b200.unet import SPECS, EngineUNet
    from oracle.unet_ref import build_unet
    torch.set_num_threads(min(8, os.cpu_count() or 8))
    oracle = build_unet(arch)       # only the seeded weights are needed on the box
    eng = EngineUNet(SPECS[arch])
    eng.load_state_dict(oracle.state_dict())
    eng.requires_grad_(False)
    eng.to("cuda")
    return oracle, eng


def _inputs(arch, n, hw, seed=3):
    import torch
    from oracle.unet_ref import CONFIGS
    cfg = CONFIGS[arch]
    g = torch.Generator().manual_seed(seed)
    x = torch.randn((n, 4, hw, hw), generator=g)
    ctx = torch.randn((n, 77, cfg.cross_attention_dim), generator=g)
    added = None
    if cfg.addition_embed_type:
        pooled = cfg.projection_class_embeddings_input_dim - 6 * cfg.addition_time_embed_dim
        added = {"text_embeds": torch.randn((n, pooled), generator=g),
                 "time_ids": torch.tensor([[hw * 8., hw * 8, 0, 0, hw * 8, hw * 8]] * n)}
    return x, ctx, added


def oracle_forward(arch, n, hw):
    """fp32 CPU oracle UNet output on bf16-rounded weights / inputs (what the engine is given)."""
    import torch
    from oracle.unet_ref import build_unet
    oracle = build_unet(arch).to(torch.bfloat16).float()
    x, ctx, added = _inputs(arch, n, hw)
    with torch.no_grad():
        return oracle(x.bfloat16().float(), torch.tensor(481), ctx.bfloat16().float(),
                      None if added is None else {k: v.bfloat16().float() for k, v in added.items()}).sample


def _make_net(unet, dev, mode=None):
    import torch
    from oracle import leco_ref
    torch.manual_seed(11)
    if mode is None:
        kw = dict(rank=4, alpha=1.0)
    elif mode == "rank32":   # stacked q|k|v rank 96 > one 64-wide K-segment: the accumulate-GEMM path
        kw = dict(rank=32, alpha=8.0)
    else:
        kw = dict(rank=4, alpha=2.0, targets=leco_ref.ATTN_TARGETS + leco_ref.CONV_TARGETS)
    with contextlib.redirect_stdout(io.StringIO()):
        net = leco_ref.LoRANetworkRef(unet, multiplier=1.0, **kw)
    g = torch.Generator().manual_seed(5)
    for l in net.unet_loras:
        l.lora_up.weight.data = (0.05 * torch.randn(l.lora_up.weight.shape, generator=g)).bfloat16().float()
        l.lora_down.weight.data = l.lora_down.weight.data.bfloat16().float()
    return net.to(dev)


def oracle_grads(arch, n, hw, mode):
    """oracle forward + autograd LoRA gradients of an MSE against a seeded goal (fp32 CPU)."""
    import torch
    from oracle.unet_ref import build_unet
    oracle = build_unet(arch).to(torch.bfloat16).float()
    x, ctx, _ = _inputs(arch, n, hw)
    net = _make_net(oracle, "cpu", mode)
    goal = torch.randn((n, 4, hw, hw), generator=torch.Generator().manual_seed(9))
    with net:
        yo = oracle(x.bfloat16().float(), torch.tensor(261), encoder_hidden_states=ctx.bfloat16().float()).sample
    lo = torch.nn.functional.mse_loss(yo, goal)
    lo.backward()
    grads = [p.grad.to(torch.bfloat16) for l in net.unet_loras for p in (l.lora_down.weight, l.lora_up.weight)]  # bf16: small fixture
    return {"y": yo.detach(), "loss": lo.item(), "grads": grads, "goal": goal}


def case_engine_forward(arch, n=2, hw=16, time_it=False):
    import torch
    from tests.oracle_cache import cached
    ref = cached(f"fwd_{arch}_{n}_{hw}", lambda: oracle_forward(arch, n, hw))
    _, eng = _engine_pair(arch)
    x, ctx, added = _inputs(arch, n, hw)
    t = torch.tensor(481)
    with torch.no_grad():
        cu_added = None if added is None else {k: v.cuda() for k, v in added.items()}
        out = eng(x.cuda(), t, encoder_hidden_states=ctx.cuda().bfloat16(), added_cond_kwargs=cu_added).sample
        torch.cuda.synchronize()
        res = _cmp(out, ref, 3e-2)
        res["rel_rms"] = ((out.float().cpu() - ref).pow(2).mean().sqrt() / ref.pow(2).mean().sqrt()).item()
        if time_it:
            xc, cc = x.cuda(), ctx.cuda().bfloat16()
            for _ in range(2):
                eng(xc, t, encoder_hidden_states=cc, added_cond_kwargs=cu_added)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.time()
            e0.record()
            for _ in range(3):
                eng(xc, t, encoder_hidden_states=cc, added_cond_kwargs=cu_added)
            e1.record()
            torch.cuda.synchronize()
            res["eager_ms_per_fwd_gpu"] = e0.elapsed_time(e1) / 3
            res["eager_ms_per_fwd_wall"] = (time.time() - t0) * 1000 / 3
    return res


def case_engine_grads(arch, n=2, hw=8, mode=None, cache_name=None):
    import torch
    from tests.oracle_cache import cached
    ora = cached(cache_name or (f"grads_{arch}" if mode is None else f"grads_{mode}_{arch}"),
                 lambda: oracle_grads(arch, n, hw, mode))
    _, eng = _engine_pair(arch)
    x, ctx, _ = _inputs(arch, n, hw)
    net_e = _make_net(eng, "cuda", mode).to(torch.bfloat16)
    with net_e:
        ye = eng(x.cuda().bfloat16(), torch.tensor(261), encoder_hidden_states=ctx.cuda().bfloat16()).sample
    le = torch.nn.functional.mse_loss(ye.float().cpu(), ora["goal"])  # loss on the CPU like train_lora.py:265-270
    le.backward()
    res = {"fwd": _cmp(ye, ora["y"], 3e-2)}
    num = den = worst = 0.0
    got = [p.grad for l in net_e.unet_loras for p in (l.lora_down.weight, l.lora_up.weight)]
    cos_min, cos_weighted, wsum = 1.0, 0.0, 0.0
    for ga, gb in zip(ora["grads"], got):
        ga, gb = ga.float(), gb.float().cpu()
        num += (ga - gb).pow(2).sum().item()
        den += ga.pow(2).sum().item()
        worst = max(worst, (ga - gb).abs().max().item() / (ga.abs().max().item() + 1e-12))
        c = (torch.dot(ga.reshape(-1), gb.reshape(-1)) / (ga.norm() * gb.norm()).clamp_min(1e-30)).item()
        cos_min = min(cos_min, c)
        cos_weighted += c * ga.pow(2).sum().item()
        wsum += ga.pow(2).sum().item()
    rel = (num / den) ** 0.5
    res["grads"] = {"rel": rel, "worst_tensor_rel": worst, "cos_min": cos_min, "cos_weighted": cos_weighted / wsum,
                    "ok": rel < 5e-2, "max_abs_err": 0, "ref_absmax": den ** 0.5}
    dl = abs(ora["loss"] - le.item()) / ora["loss"]
    res["loss"] = {"rel": dl, "ok": dl < 2e-2, "max_abs_err": 0, "ref_absmax": ora["loss"]}
    return _merge(res)


def case_text_kernels():
    """leco_embed_tokens / leco_activation / leco_softmax_rows_causal vs plain torch (tests/torch_backend.py)."""
    import torch
    from leco_b200 import ops
    from tests import torch_backend as tb
    res = {}
    tok, pos = _rand((600, 128), 0.1, 1), _rand((77, 128), 0.1, 2)
    ids = torch.randint(0, 600, (3 * 77,), generator=torch.Generator().manual_seed(3)).int().cuda()
    out = ops.embed_tokens(ids, tok, pos, 77)
    ref = tb.embed_tokens(ids.cpu(), tok.cpu(), pos.cpu(), 77)
    res["embed_tokens"] = _cmp(out, ref, tol=1e-6)                       # one rounding of an exact fp32 sum
    res["embed_tokens"]["ok"] = bool(torch.equal(out.cpu(), ref))
    x = _rand((231, 3072), 2.0, 4)
    for kind, name in ((0, "silu"), (1, "quick_gelu"), (2, "gelu")):
        res[f"act_{name}"] = _cmp(ops.activation(x, kind), tb.activation(x.cpu().float(), kind), tol=8e-3)
    for (b, h, sq, pad) in ((2, 12, 77, 80), (1, 4, 64, 64), (3, 2, 5, 16)):
        sc = (_rand((b, h, sq, pad), 3.0, 5, dtype=torch.float32)).contiguous()
        p = ops.softmax_rows(sc, sq, pad, causal_sq=sq)
        pr = tb.softmax_rows(sc.cpu(), sq, pad, causal_sq=sq)
        res[f"softmax_causal_{sq}"] = _cmp(p, pr, tol=8e-3)
        upper = torch.ones(sq, pad, dtype=torch.bool).triu(1)
        res[f"softmax_causal_{sq}"]["ok"] &= bool((p.cpu().float()[..., upper] == 0).all())   # the mask is exact
        o, _ = ops.attention_v0(*( _rand((b * sq, h * 64), 1.0, 6 + i) for i in range(3)), b, sq, sq, h, 64, 0.125, causal=True)
        orf, _ = tb.attention_v0(*( _rand((b * sq, h * 64), 1.0, 6 + i).cpu() for i in range(3)), b, sq, sq, h, 64, 0.125, causal=True)
        res[f"attn_causal_{sq}"] = _cmp(o, orf)
    return _merge(res)


def case_text_encoder(name, batch=3, golden=False):
    """Engine CLIP text encoder vs the fp32 CPU oracle (oracle/clip_ref.py; golden=True: vs transformers' own outputs in
    tests/golden/clip_tiny.pt).  Bound: no farther from fp32 than 1.5x a stock-torch bf16 run of the oracle + 0.5 %."""
    import torch
    from leco_b200.text_encoder import TEXT_SPECS, ClipTextEncoder, build_text_encoder
    from oracle import clip_ref
    from tests.clip_fixtures import token_ids_for
    spec = TEXT_SPECS[name]
    if golden:
        blob = torch.load(os.path.join(ROOT, "tests", "golden", "clip_tiny.pt"), weights_only=False)[name]
        enc = ClipTextEncoder(spec)
        enc.load_state_dict({k: v.float() for k, v in blob["state_dict"].items()})
        enc = enc.cuda()
        ids = blob["ids"]
    else:
        enc = build_text_encoder(name, device="cpu", seed=1)
        with torch.no_grad():
            for p_ in enc.parameters():
                p_.copy_(p_.to(torch.bfloat16).float())           # both sides see bf16-exact weights
        enc = enc.cuda()
        ids = token_ids_for(spec, batch, seed=2)
    sd = {k: v.detach().float().cpu() for k, v in enc.state_dict().items()}
    kw = dict(heads=spec.num_attention_heads, act=spec.hidden_act, eps=spec.layer_norm_eps, eos_token_id=spec.eos_token_id)
    last, pooled, emb, hidden = clip_ref.clip_text_forward(sd, ids, **kw)
    if golden:
        assert (last - blob["last_hidden_state"]).abs().max() < 2e-5
    sd_gpu = {k: v.cuda() for k, v in sd.items()}
    y_last, _, y_emb, y_hidden = clip_ref.clip_text_forward(sd_gpu, ids.cuda(), dtype=torch.bfloat16, **kw)   # yardstick
    t0 = time.time()
    out = enc(ids, output_hidden_states=True)
    torch.cuda.synchronize()
    res = {}

    def bound(got, want, yard, key):
        want = want.float()
        den = want.pow(2).mean().sqrt().item() + 1e-12
        e = (got.float().cpu() - want).pow(2).mean().sqrt().item() / den
        ey = (yard.float().cpu() - want).pow(2).mean().sqrt().item() / den
        res[key] = {"rel": e, "bf16_torch_rel": ey, "ok": bool(torch.isfinite(got.float()).all()) and e <= 1.5 * ey + 5e-3,
                    "max_abs_err": (got.float().cpu() - want).abs().max().item(), "ref_absmax": want.abs().max().item()}
    bound(out.last_hidden_state, last, y_last, "last_hidden_state")
    bound(out.hidden_states[-2], hidden[-2], y_hidden[-2], "penultimate")
    if spec.projection_dim:
        bound(out[0], emb, y_emb, "text_embeds")
        res["first_is_text_embeds"] = {"rel": 0.0, "ok": out[0].shape == (ids.shape[0], spec.projection_dim)}
    else:
        res["first_is_last"] = {"rel": 0.0, "ok": out[0] is out.last_hidden_state and out[0].shape == last.shape}
    r = _merge(res)
    r["encode_s"] = round(time.time() - t0, 3)
    return r


CASES = [
    ("text_kernels", case_text_kernels, {}),
    ("text_encoder_tiny_golden", case_text_encoder, dict(name="tiny_clip", golden=True)),
    ("text_encoder_tiny_proj_golden", case_text_encoder, dict(name="tiny_clip_proj", golden=True)),
    ("text_encoder_clip_l", case_text_encoder, dict(name="clip_l", batch=4)),
    ("text_encoder_openclip_h", case_text_encoder, dict(name="openclip_h", batch=2)),
    ("text_encoder_openclip_bigg", case_text_encoder, dict(name="openclip_bigg", batch=2)),
    ("norms", case_norms, {}),
    ("elementwise", case_elementwise, {}),
    ("training_kernels", case_training_kernels, {}),
    ("optimizers", case_optimizers, {}),
    ("optimizers_master", case_optimizers_master, {}),
    ("sched_step", case_sched_step, {}),
    ("transpose_tiles", case_transpose_tiles, {}),
    ("determinism", case_determinism, {}),
    ("attn_self_4096_d64", case_attention, dict(nb=1, sq=4096, skv=4096, heads=5, d=64)),
    ("attn_self_4096_d64_v0", case_attention, dict(nb=1, sq=4096, skv=4096, heads=5, d=64, bwd_impl="v0")),
    ("attn_self_1024_d64", case_attention, dict(nb=2, sq=1024, skv=1024, heads=5, d=64)),
    ("attn_self_1024_d64_v0", case_attention, dict(nb=2, sq=1024, skv=1024, heads=5, d=64, bwd_impl="v0")),
    ("attn_cross_77_v0", case_attention, dict(nb=2, sq=256, skv=77, heads=2, d=64, cross=True, bwd_impl="v0")),
    ("attn_cross_77_sq4096", case_attention, dict(nb=1, sq=4096, skv=77, heads=5, d=64, cross=True)),
    ("attn_self_300_ragged", case_attention, dict(nb=2, sq=300, skv=300, heads=3, d=64)),
    ("attn_self_d160_v0", case_attention, dict(nb=1, sq=256, skv=256, heads=2, d=160)),
    ("attn_cross_77", case_attention, dict(nb=2, sq=256, skv=77, heads=2, d=64, cross=True)),
    ("attn_self_d40", case_attention, dict(nb=2, sq=256, skv=256, heads=8, d=40)),
    ("attn_self_64_d8", case_attention, dict(nb=2, sq=64, skv=64, heads=8, d=8)),
    ("attn_self_tiny_sq4", case_attention, dict(nb=2, sq=4, skv=4, heads=4, d=64)),
    ("flash_self_1024_m0", case_flash, dict(nb=2, sq=1024, skv=1024, heads=5, d=64, v_mode=0)),
    ("flash_self_1024_m1", case_flash, dict(nb=2, sq=1024, skv=1024, heads=5, d=64, v_mode=1)),
    ("flash_cross_77_m0", case_flash, dict(nb=2, sq=256, skv=77, heads=2, d=64, cross=True, v_mode=0)),
    ("flash_cross_77_m1", case_flash, dict(nb=2, sq=256, skv=77, heads=2, d=64, cross=True, v_mode=1)),
    ("flash_d40_m0", case_flash, dict(nb=2, sq=256, skv=256, heads=8, d=40, v_mode=0)),
    ("flash_d40_m1", case_flash, dict(nb=2, sq=256, skv=256, heads=8, d=40, v_mode=1)),
    ("flash_sq64_m0", case_flash, dict(nb=3, sq=64, skv=64, heads=4, d=64, v_mode=0)),
    ("flash_sq200_skv300_m1", case_flash, dict(nb=1, sq=200, skv=300, heads=2, d=64, cross=True, v_mode=1)),
    ("flash_sq200_skv300_m0", case_flash, dict(nb=1, sq=200, skv=300, heads=2, d=64, cross=True, v_mode=0)),
    ("flash_rescale_2048", case_flash, dict(nb=2, sq=2048, skv=2048, heads=4, d=64, qscale=8.0)),
    ("flash_rescale_ramp_1000", case_flash, dict(nb=2, sq=1000, skv=1000, heads=4, d=64, qscale=3.0, kramp=6.0, v_mode=1)),
    ("flash_rescale_2048_smemP", case_flash, dict(nb=2, sq=2048, skv=2048, heads=4, d=64, qscale=8.0, ts=0)),
    ("flash_self_1024_m0_smemP", case_flash, dict(nb=2, sq=1024, skv=1024, heads=5, d=64, v_mode=0, ts=0)),
    ("flash_cross_77_m1_smemP", case_flash, dict(nb=2, sq=256, skv=77, heads=2, d=64, cross=True, v_mode=1, ts=0)),
    ("flash_perf_4096_m0_smemP", case_flash, dict(nb=4, sq=4096, skv=4096, heads=5, d=64, v_mode=0, perf=True, ts=0)),
    ("flash_perf_4096_m0", case_flash, dict(nb=4, sq=4096, skv=4096, heads=5, d=64, v_mode=0, perf=True)),
    ("flash_perf_4096_m1", case_flash, dict(nb=4, sq=4096, skv=4096, heads=5, d=64, v_mode=1, perf=True)),
    # head dims above 64 (SD1.5: 8 heads at every level -> 80 at 32x32, 160 at 16x16 / 8x8; cross-attention Skv = 77)
    ("flash_wide_d80_1024", case_flash, dict(nb=2, sq=1024, skv=1024, heads=8, d=80)),
    ("flash_wide_d80_cross77", case_flash, dict(nb=2, sq=1024, skv=77, heads=8, d=80, cross=True)),
    ("flash_wide_d160_256", case_flash, dict(nb=2, sq=256, skv=256, heads=8, d=160)),
    ("flash_wide_d160_64", case_flash, dict(nb=2, sq=64, skv=64, heads=8, d=160)),
    ("flash_wide_d160_cross77", case_flash, dict(nb=2, sq=256, skv=77, heads=8, d=160, cross=True)),
    ("flash_wide_d128_sq200_skv300", case_flash, dict(nb=1, sq=200, skv=300, heads=2, d=128, cross=True)),
    ("flash_wide_d96_rescale_1024", case_flash, dict(nb=2, sq=1024, skv=1024, heads=4, d=96, qscale=8.0)),
    ("flash_wide_d192_576", case_flash, dict(nb=1, sq=576, skv=576, heads=3, d=192)),
    ("flash_wide_d72_ramp_1000", case_flash, dict(nb=1, sq=1000, skv=1000, heads=2, d=72, qscale=3.0, kramp=6.0)),
    ("flash_perf_wide_d80_1024", case_flash, dict(nb=8, sq=1024, skv=1024, heads=8, d=80, perf=True)),
    ("flash_perf_wide_d160_256", case_flash, dict(nb=8, sq=256, skv=256, heads=8, d=160, perf=True)),
    ("engine_fwd_tiny21", case_engine_forward, dict(arch="tiny21")),
    ("engine_fwd_tiny15", case_engine_forward, dict(arch="tiny15")),
    ("engine_fwd_tinyxl", case_engine_forward, dict(arch="tinyxl")),
    ("engine_grads_tiny21", case_engine_grads, dict(arch="tiny21")),
    ("engine_grads_tiny15", case_engine_grads, dict(arch="tiny15")),
    ("engine_grads_c3lier_tiny15", case_engine_grads, dict(arch="tiny15", mode="c3lier")),
    ("engine_grads_rank32_tiny21", case_engine_grads, dict(arch="tiny21", mode="rank32")),
    ("engine_fwd_sd21_64", case_engine_forward, dict(arch="sd21", n=2, hw=64, time_it=True)),
    ("perf_norms", case_norm_perf, {}),
]


def main():
    if len(sys.argv) >= 3 and sys.argv[1] == "--case":
        for n, fn, kw in CASES:
            if n == sys.argv[2]:
                print("RESULT " + json.dumps(fn(**kw)))
                return
        raise SystemExit("unknown case")
    only = sys.argv[1:] or None
    out_dir = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out_dir, exist_ok=True)
    results = {}
    for name, _, kw in CASES:
        if only and not any(name.startswith(o) for o in only):
            continue
        t0 = time.time()
        try:
            pr = subprocess.run([sys.executable, os.path.abspath(__file__), "--case", name], capture_output=True,
                                text=True, timeout=150)
            line = [l for l in pr.stdout.splitlines() if l.startswith("RESULT ")]
            res = json.loads(line[-1][7:]) if pr.returncode == 0 and line else {
                "ok": False, "rc": pr.returncode, "stderr": pr.stderr[-2500:], "stdout": pr.stdout[-800:]}
        except subprocess.TimeoutExpired:
            res = {"ok": False, "timeout": True}
        res["secs"] = round(time.time() - t0, 1)
        results[name] = res
        brief = {k: v for k, v in res.items() if k != "parts"}
        if not res.get("ok") and "parts" in res:
            brief["failed_parts"] = {k: v for k, v in res["parts"].items() if not v.get("ok")}
        print(name, json.dumps(brief)[:1500], flush=True)
        with open(os.path.join(out_dir, "kernel_cases.json"), "w") as f:
            json.dump(results, f, indent=1)
    print(f"SUMMARY {sum(1 for r in results.values() if r.get('ok'))}/{len(results)} ok")


if __name__ == "__main__":
    main()
