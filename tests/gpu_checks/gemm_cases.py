"""GPU check harness for leco_gemm_bf16 (run on the B200 box through gpurun).

Every case runs in its own subprocess under a timeout so that a trap / hang in one case
cannot take the others down.  Results -> gpurun_out/gemm_cases.json.  The comparison is a
plain PyTorch fp32 reference of the same op on the same bf16-rounded inputs.
"""
from __future__ import annotations

import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def _rand(shape, scale=1.0, seed=0):
    import torch
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).to(torch.bfloat16).cuda()


def _cmp(out, ref, tol=2e-2):
    import torch
    out = out.float()
    err = (out - ref).abs().max().item()
    scale = ref.abs().max().item() + 1e-6
    finite = bool(torch.isfinite(out).all().item())
    return {"max_abs_err": err, "ref_absmax": scale, "rel": err / scale, "ok": finite and err / scale < tol}


def case_matrix(M, N, K, block_n=0, bias=False, rowbias=0, residual=False, lora=0, geglu=False,
                alpha=1.0, out_fp32=False, strided=False, cta_pair=0):
    import torch
    from leco_b200 import ops
    a = _rand((M, K), seed=1)
    if strided:  # A is a column slice of a wider buffer (like Q inside a fused QKV output)
        wide = _rand((M, K * 3), seed=1)
        a = wide[:, K:2 * K]
    b = _rand((N, K), scale=K ** -0.5, seed=2)
    ref = a.float() @ b.float().t()
    kw = {}
    if lora:
        t = _rand((M, lora), seed=3)
        up = _rand((N, lora), scale=0.1, seed=4)
        ref = ref + t.float() @ up.float().t()
        kw.update(lora_t=t, lora_up=up)
    ref = ref * alpha
    if bias:
        bv = _rand((N,), seed=5)
        ref = ref + bv.float()[None]
        kw["bias"] = bv
    if rowbias:
        groups = (M + rowbias - 1) // rowbias
        rb = _rand((groups, N), seed=6)
        ref = ref + rb.float().repeat_interleave(rowbias, 0)[:M]
        kw.update(rowbias=rb, rows_per_group=rowbias)
    if geglu:
        h, gte = ref[:, :N // 2], ref[:, N // 2:]
        ref = h * torch.nn.functional.gelu(gte)
    if residual:
        r = _rand(ref.shape, seed=7)
        ref = ref + r.float()
        kw["residual"] = r
    out = ops.gemm(a, b, geglu=geglu, alpha=alpha, out_fp32=out_fp32, block_n=block_n, cta_pair=cta_pair, **kw)
    torch.cuda.synchronize()
    return _cmp(out, ref)


def case_batched(B1, B0, M, N, K, out_fp32=True, alpha=0.125):
    """S = alpha * Q K^T with Q,K strided views into a fused [B1*M, 3*B0*K] projection output."""
    import torch
    from leco_b200 import ops
    C = B0 * K
    qkv = _rand((B1 * max(M, N), 3 * C), seed=11)
    q = qkv[: B1 * M, 0:C].unflatten(0, (B1, M)).unflatten(2, (B0, K)).permute(0, 2, 1, 3)
    k = qkv[: B1 * N, C:2 * C].unflatten(0, (B1, N)).unflatten(2, (B0, K)).permute(0, 2, 1, 3)
    out = torch.empty((B1, B0, M, N), device="cuda", dtype=torch.float32 if out_fp32 else torch.bfloat16)
    ops.gemm_batched(q, k, out, alpha=alpha)
    torch.cuda.synchronize()
    ref = alpha * torch.einsum("bhmk,bhnk->bhmn", q.float(), k.float())
    return _cmp(out, ref)


def case_conv(n, h, w, cin, cout, block_n=0, bias=True, rowbias=True, residual=True, lora=0, cta_pair=0):
    import torch
    import torch.nn.functional as F
    from leco_b200 import ops
    x = _rand((n, h, w, cin), seed=21)                       # NHWC
    wt = _rand((cout, cin, 3, 3), scale=(9 * cin) ** -0.5, seed=22)  # OIHW
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), wt.float(), padding=1).permute(0, 2, 3, 1).reshape(n * h * w, cout)
    wk = wt.permute(0, 2, 3, 1).reshape(cout, 9 * cin).contiguous()  # [O, (kh,kw,c)]
    kw = {}
    if lora:
        t = _rand((n * h * w, lora), seed=23)
        up = _rand((cout, lora), scale=0.1, seed=24)
        ref = ref + t.float() @ up.float().t()
        kw.update(lora_t=t, lora_up=up)
    if bias:
        bv = _rand((cout,), seed=25)
        ref = ref + bv.float()[None]
        kw["bias"] = bv
    if rowbias:
        rb = _rand((n, cout), seed=26)
        ref = ref + rb.float().repeat_interleave(h * w, 0)
        kw.update(rowbias=rb, rows_per_group=h * w)
    if residual:
        r = _rand(ref.shape, seed=27)
        ref = ref + r.float()
        kw["residual"] = r
    out = ops.gemm(x.reshape(n * h * w, cin), wk, conv_nhw=(n, h, w), block_n=block_n, cta_pair=cta_pair, **kw)
    torch.cuda.synchronize()
    return _cmp(out, ref)


def case_fl(M, N, K, kl=16, rank=4, geglu=False, conv=None, bias=True, residual=True, t_out=False, block_n=0):
    """In-kernel LoRA: out = a W^T + s (a Ad^T) Bup^T (+ epilogue), optional saved T."""
    import torch
    import torch.nn.functional as F
    from leco_b200 import ops
    sc = 0.25
    if conv:
        n, h, w, cin = conv
        M, K = n * h * w, 9 * cin
        x = _rand((n, h, w, cin), seed=21)
        wt = _rand((N, cin, 3, 3), scale=(9 * cin) ** -0.5, seed=22)
        a = x.reshape(M, cin)
        wk = wt.permute(0, 2, 3, 1).reshape(N, K).contiguous()
        adw = torch.zeros((kl, cin, 3, 3))
        adw[:rank] = torch.randn((rank, cin, 3, 3), generator=torch.Generator().manual_seed(5)) * (9 * cin) ** -0.5
        adw = adw.to(torch.bfloat16).cuda()
        ad = adw.permute(0, 2, 3, 1).reshape(kl, K).contiguous()
        base = F.conv2d(x.float().permute(0, 3, 1, 2), wt.float(), padding=1).permute(0, 2, 3, 1).reshape(M, N)
        t = sc * F.conv2d(x.float().permute(0, 3, 1, 2), adw.float(), padding=1).permute(0, 2, 3, 1).reshape(M, kl)
        kw = dict(conv_nhw=(n, h, w))
    else:
        a = _rand((M, K), seed=1)
        wk = _rand((N, K), scale=K ** -0.5, seed=2)
        ad = torch.zeros((kl, K))
        ad[:rank] = torch.randn((rank, K), generator=torch.Generator().manual_seed(5)) * K ** -0.5
        ad = ad.to(torch.bfloat16).cuda()
        base = a.float() @ wk.float().t()
        t = sc * (a.float() @ ad.float().t())
        kw = {}
    bup = torch.zeros((N, kl))
    bup[:, :rank] = torch.randn((N, rank), generator=torch.Generator().manual_seed(6)) * 0.3
    bup = bup.to(torch.bfloat16).cuda()
    ref = base + t @ bup.float().t()
    if bias:
        bv = _rand((N,), seed=7)
        ref = ref + bv.float()[None]
        kw["bias"] = bv
    if geglu:
        hh, gg = ref[:, :N // 2], ref[:, N // 2:]
        ref = hh * F.gelu(gg)
    if residual and not geglu:
        r = _rand(ref.shape, seed=8)
        ref = ref + r.float()
        kw["residual"] = r
    tbuf = torch.zeros((M, kl), device="cuda", dtype=torch.bfloat16) if t_out else None
    out = ops.gemm(a, wk, geglu=geglu, fl_ad=ad, fl_bup=bup, fl_scale=sc, fl_rank=rank, fl_t_out=tbuf, block_n=block_n, **kw)
    torch.cuda.synchronize()
    res = _cmp(out, ref)
    if t_out:
        rt = _cmp(tbuf, t)
        res["t_rel"] = rt["rel"]
        res["ok"] = res["ok"] and rt["ok"]
    return res


def case_mma_rate():
    """Perf triage: one 128 x BN tile per SM, K = 16384 (256 chunks), for every BN and pipeline mode
    (0 normal, 1 TMA only, 2 MMA only).  Reports microseconds per K-chunk."""
    import torch
    from leco_b200 import ops
    out = {"ok": True}
    K = 16384
    for bn in (64, 128, 160, 256):
        a = _rand((148 * 128, K), seed=1)
        b = _rand((bn, K), scale=K ** -0.5, seed=2)
        o = torch.empty((148 * 128, bn), device="cuda", dtype=torch.bfloat16)
        for mode in (0, 1, 2):
            ops._GEMM_DEBUG_MODE = mode
            for _ in range(2):
                ops.gemm(a, b, o, block_n=bn)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                ops.gemm(a, b, o, block_n=bn)
            e1.record()
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) / 5 * 1e3
            out[f"bn{bn}_mode{mode}_us_per_chunk"] = round(us / (K // 64), 4)
        ops._GEMM_DEBUG_MODE = 0
    return out


def case_fl_perf():
    """in-kernel LoRA vs separate T GEMM + K-segment on the UNet's level-0 shapes."""
    import torch
    from leco_b200 import ops
    res = {"ok": True}
    for (M, N, K, geglu) in ((16384, 320, 320, False), (16384, 960, 320, False), (16384, 2560, 320, True),
                             (16384, 320, 1280, False), (4096, 640, 640, False)):
        a = _rand((M, K), seed=1)
        w = _rand((N, K), scale=K ** -0.5, seed=2)
        ad = _rand((16, K), scale=K ** -0.5, seed=3)
        bup = _rand((N, 16), scale=0.1, seed=4)
        o = torch.empty((M, N // 2 if geglu else N), device="cuda", dtype=torch.bfloat16)

        def t_plain():
            ops.gemm(a, w, o, geglu=geglu)

        def t_seg():
            t = ops.gemm(a, ad, alpha=0.25)
            ops.gemm(a, w, o, geglu=geglu, lora_t=t, lora_up=bup)

        def t_fl():
            ops.gemm(a, w, o, geglu=geglu, fl_ad=ad, fl_bup=bup, fl_scale=0.25, fl_rank=4)
        for name, fn in (("plain", t_plain), ("seg2", t_seg), ("fused", t_fl)):
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                fn()
            e1.record()
            torch.cuda.synchronize()
            res[f"M{M}_N{N}_K{K}_{name}_us"] = round(e0.elapsed_time(e1) / 20 * 1e3, 2)
    return res


def case_perf(M, N, K, block_n=0, conv=None, iters=20, cta_pair=0, graph=False):
    import torch
    from leco_b200 import ops
    if conv:
        n, h, w, cin = conv
        a = _rand((n * h * w, cin), seed=31)
        K = 9 * cin
        M = n * h * w
    else:
        a = _rand((M, K), seed=31)
    b = _rand((N, K), scale=K ** -0.5, seed=32)
    out = torch.empty((M, N), device="cuda", dtype=torch.bfloat16)
    kw = dict(conv_nhw=conv[:3]) if conv else {}
    kw['cta_pair'] = cta_pair
    for _ in range(3):
        ops.gemm(a, b, out, block_n=block_n, **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if graph:
        # device time without the host's per-launch cost: replay `iters` captured launches (as the trainer does)
        g = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            ops.gemm(a, b, out, block_n=block_n, **kw)
        torch.cuda.current_stream().wait_stream(side)
        with torch.cuda.graph(g):
            for _ in range(iters):
                ops.gemm(a, b, out, block_n=block_n, **kw)
        g.replay()
        torch.cuda.synchronize()
        e0.record()
        g.replay()
        e1.record()
    else:
        e0.record()
        for _ in range(iters):
            ops.gemm(a, b, out, block_n=block_n, **kw)
        e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    tf = 2.0 * M * N * K / (ms * 1e-3) / 1e12
    # cuBLAS yardstick (reference only, never on the product path)
    if not conv:
        for _ in range(3):
            torch.matmul(a, b.t())
        torch.cuda.synchronize()
        e0.record()
        for _ in range(iters):
            torch.matmul(a, b.t())
        e1.record()
        torch.cuda.synchronize()
        ms_ref = e0.elapsed_time(e1) / iters
    else:
        ms_ref = None
    return {"ok": True, "ms": ms, "tflops": tf, "cublas_ms": ms_ref}


def case_shape_modes():
    """The UNet's dominant GEMM shapes timed under the pipeline triage modes:
    0 normal, 1 TMA only, 2 MMA only, 3 no epilogue.  Microseconds per launch."""
    from leco_b200 import ops
    res = {"ok": True}
    shapes = {
        "conv64_320": dict(M=0, N=320, K=0, conv=(4, 64, 64, 320)),
        "conv32_640": dict(M=0, N=640, K=0, conv=(4, 32, 32, 640)),
        "conv16_1280": dict(M=0, N=1280, K=0, conv=(4, 16, 16, 1280)),
        "conv64_640to320": dict(M=0, N=320, K=0, conv=(4, 64, 64, 640)),
        "ff1": dict(M=16384, N=2560, K=320),
        "ff2": dict(M=16384, N=320, K=1280),
        "qkv320": dict(M=16384, N=320, K=320),
        "lin640": dict(M=4096, N=640, K=640),
    }
    for name, kw in shapes.items():
        for mode in (0, 1, 2, 3):
            ops._GEMM_DEBUG_MODE = mode
            try:
                r = case_perf(iters=20, graph=True, **kw)
            finally:
                ops._GEMM_DEBUG_MODE = 0
            res[f"{name}_mode{mode}_us"] = round(r["ms"] * 1e3, 2)
            if mode == 0 and "conv" in kw or mode == 0 and kw.get("M", 0) >= 4096:
                ops._GEMM_DEBUG_MODE = 0
                r2 = case_perf(iters=20, graph=True, cta_pair=1, **kw)
                res[f"{name}_2cta_us"] = round(r2["ms"] * 1e3, 2)
            if mode == 0:
                res[f"{name}_tflops"] = round(r["tflops"], 1)
                if r["cublas_ms"]:
                    res[f"{name}_cublas_us"] = round(r["cublas_ms"] * 1e3, 2)
    return res


CASES = [
    ("basic_bn128", case_matrix, dict(M=256, N=256, K=256, block_n=128)),
    ("basic_bn64", case_matrix, dict(M=256, N=256, K=256, block_n=64)),
    ("basic_bn160", case_matrix, dict(M=256, N=320, K=256, block_n=160)),
    ("basic_bn256", case_matrix, dict(M=256, N=512, K=256, block_n=256)),
    ("multi_tile_persist", case_matrix, dict(M=4096, N=2560, K=320)),
    ("ragged_m308_n320", case_matrix, dict(M=308, N=320, K=1024)),
    ("tiny_m4", case_matrix, dict(M=4, N=1280, K=320, bias=True)),
    ("k80", case_matrix, dict(M=256, N=128, K=80)),
    ("k48", case_matrix, dict(M=128, N=64, K=48)),
    ("n_ragged_n72", case_matrix, dict(M=256, N=72, K=128, block_n=64)),
    ("bias_rowbias_res", case_matrix, dict(M=512, N=320, K=320, bias=True, rowbias=128, residual=True)),
    ("lora16", case_matrix, dict(M=1024, N=960, K=320, lora=16)),
    ("lora32_bias_res", case_matrix, dict(M=1024, N=640, K=640, lora=32, bias=True, residual=True)),
    ("lora_only_small_n16", case_matrix, dict(M=1024, N=16, K=320, alpha=0.25)),
    ("geglu", case_matrix, dict(M=512, N=2560, K=320, geglu=True, bias=True)),
    ("geglu_lora", case_matrix, dict(M=300, N=1024, K=128, geglu=True, bias=True, lora=16)),
    ("fp32_out_alpha", case_matrix, dict(M=256, N=256, K=64, alpha=0.125, out_fp32=True)),
    ("strided_a", case_matrix, dict(M=512, N=128, K=128, strided=True)),
    ("batched_qk_d64", case_batched, dict(B1=2, B0=5, M=1024, N=1024, K=64)),
    ("batched_qk_cross77", case_batched, dict(B1=4, B0=5, M=256, N=80, K=64)),
    ("batched_bf16out", case_batched, dict(B1=2, B0=2, M=128, N=128, K=64, out_fp32=False)),
    ("conv_64x64_c64", case_conv, dict(n=2, h=64, w=64, cin=64, cout=64)),
    ("conv_32x32_c128_n320", case_conv, dict(n=2, h=32, w=32, cin=128, cout=320)),
    ("conv_16x16", case_conv, dict(n=4, h=16, w=16, cin=192, cout=128)),
    ("conv_8x8_n3", case_conv, dict(n=3, h=8, w=8, cin=128, cout=256)),
    ("conv_4x4", case_conv, dict(n=4, h=4, w=4, cin=64, cout=64)),
    ("conv_2x2", case_conv, dict(n=4, h=2, w=2, cin=64, cout=64)),
    ("conv_40x40", case_conv, dict(n=2, h=40, w=40, cin=64, cout=64)),
    ("conv_24x40_rect", case_conv, dict(n=1, h=24, w=40, cin=64, cout=64)),
    ("conv_lora16", case_conv, dict(n=2, h=32, w=32, cin=128, cout=128, lora=16)),
    ("conv_plain", case_conv, dict(n=2, h=32, w=32, cin=64, cout=64, bias=False, rowbias=False, residual=False)),
    # few-tile long-K problems: split-K with the finalize done by the last-arriving K-slice CTA (no finalize launch)
    ("splitk_m256_bias_rb_res", case_matrix, dict(M=256, N=1280, K=5120, bias=True, rowbias=64, residual=True)),
    ("splitk_m308_ragged", case_matrix, dict(M=308, N=1288, K=4096, bias=True, residual=True)),
    ("splitk_conv_16_1280", case_conv, dict(n=4, h=16, w=16, cin=1280, cout=1280)),
    ("splitk_conv_8_1280", case_conv, dict(n=4, h=8, w=8, cin=1280, cout=1280)),
    ("triage_mma_rate", case_mma_rate, dict()),
    ("triage_fl_perf", case_fl_perf, dict()),
    ("triage_shape_modes", case_shape_modes, dict()),
    ("fl_linear_r4", case_fl, dict(M=1024, N=320, K=320)),
    ("fl_linear_qkv_r12_tout", case_fl, dict(M=4096, N=960, K=320, rank=12, t_out=True)),
    ("fl_linear_kl32_bn128", case_fl, dict(M=512, N=640, K=640, kl=32, rank=24, block_n=128, t_out=True)),
    ("fl_linear_bn64_ragged", case_fl, dict(M=300, N=72, K=128, block_n=64)),
    # many tiles per CTA: runs of N tiles that share one staged T (the fresh / reuse logic of the N-inner tile walk)
    ("fl_runs_qkv_16384", case_fl, dict(M=16384, N=960, K=320, rank=12, t_out=True)),
    ("fl_runs_geglu_4096", case_fl, dict(M=4096, N=2560, K=320, geglu=True)),
    ("fl_runs_ragged_m", case_fl, dict(M=5000, N=640, K=640, kl=32, rank=20, t_out=True)),
    ("fl_geglu", case_fl, dict(M=512, N=2560, K=320, geglu=True)),
    ("fl_splitk_m256", case_fl, dict(M=256, N=1280, K=5120)),
    ("fl_ctx_m308", case_fl, dict(M=308, N=640, K=1024, rank=8)),
    ("fl_conv_32", case_fl, dict(M=0, N=128, K=0, conv=(2, 32, 32, 128), t_out=True)),
    ("fl_conv_8_splitk", case_fl, dict(M=0, N=1280, K=0, conv=(4, 8, 8, 1280), rank=8)),
    ("splitk_conv16_1280", case_conv, dict(n=4, h=16, w=16, cin=1280, cout=1280)),
    ("splitk_conv8_lora", case_conv, dict(n=4, h=8, w=8, cin=1280, cout=1280, lora=16)),
    ("splitk_linear_m308", case_matrix, dict(M=308, N=2560, K=1024, bias=True)),
    ("splitk_linear_m256_res", case_matrix, dict(M=256, N=1280, K=5120, bias=True, residual=True, lora=16)),
    ("perf_conv_8_1280", case_perf, dict(M=0, N=1280, K=0, conv=(4, 8, 8, 1280))),
    ("2cta_basic_bn128", case_matrix, dict(M=512, N=256, K=256, block_n=128, cta_pair=1)),
    ("2cta_basic_bn256", case_matrix, dict(M=512, N=512, K=512, block_n=256, cta_pair=1)),
    ("2cta_basic_bn160", case_matrix, dict(M=256, N=320, K=256, block_n=160, cta_pair=1)),
    ("2cta_basic_bn64", case_matrix, dict(M=256, N=128, K=128, block_n=64, cta_pair=1)),
    ("2cta_odd_tiles_m384", case_matrix, dict(M=384, N=256, K=320, cta_pair=1)),
    ("2cta_ragged_m308", case_matrix, dict(M=308, N=320, K=1024, cta_pair=1)),
    ("2cta_tiny_m4", case_matrix, dict(M=4, N=1280, K=320, bias=True, cta_pair=1)),
    ("2cta_persist", case_matrix, dict(M=16384, N=2560, K=320, cta_pair=1)),
    ("2cta_bias_rowbias_res", case_matrix, dict(M=512, N=320, K=320, bias=True, rowbias=128, residual=True, cta_pair=1)),
    ("2cta_lora16", case_matrix, dict(M=1024, N=960, K=320, lora=16, cta_pair=1)),
    ("2cta_geglu_lora", case_matrix, dict(M=300, N=1024, K=128, geglu=True, bias=True, lora=16, cta_pair=1)),
    ("2cta_k80", case_matrix, dict(M=256, N=128, K=80, cta_pair=1)),
    ("2cta_conv_64x64", case_conv, dict(n=2, h=64, w=64, cin=64, cout=64, cta_pair=1)),
    ("2cta_conv_32_c128_n320", case_conv, dict(n=2, h=32, w=32, cin=128, cout=320, cta_pair=1)),
    ("2cta_conv_8x8_n3", case_conv, dict(n=3, h=8, w=8, cin=128, cout=256, cta_pair=1)),
    ("2cta_conv_2x2", case_conv, dict(n=4, h=2, w=2, cin=64, cout=64, cta_pair=1)),
    ("2cta_conv_40x40", case_conv, dict(n=2, h=40, w=40, cin=64, cout=64, cta_pair=1)),
    ("2cta_conv_lora16", case_conv, dict(n=2, h=32, w=32, cin=128, cout=128, lora=16, cta_pair=1)),
    ("perf2_ff1", case_perf, dict(M=16384, N=2560, K=320, cta_pair=1)),
    ("perf2_ff2", case_perf, dict(M=16384, N=320, K=1280, cta_pair=1)),
    ("perf2_8k", case_perf, dict(M=8192, N=8192, K=8192, iters=5, cta_pair=1)),
    ("perf2_conv_64_320", case_perf, dict(M=0, N=320, K=0, conv=(4, 64, 64, 320), cta_pair=1)),
    ("perf2_conv_32_640", case_perf, dict(M=0, N=640, K=0, conv=(4, 32, 32, 640), cta_pair=1)),
    ("perf2_conv_16_1280", case_perf, dict(M=0, N=1280, K=0, conv=(4, 16, 16, 1280), cta_pair=1)),
    ("perf_conv_32_640", case_perf, dict(M=0, N=640, K=0, conv=(4, 32, 32, 640))),
    ("perf_ff1", case_perf, dict(M=16384, N=2560, K=320)),
    ("perf_ff2", case_perf, dict(M=16384, N=320, K=1280)),
    ("perf_8k", case_perf, dict(M=8192, N=8192, K=8192, iters=5)),
    ("perf_8k_bn128", case_perf, dict(M=8192, N=8192, K=8192, block_n=128, iters=5)),
    ("perf_conv_64_320", case_perf, dict(M=0, N=320, K=0, conv=(4, 64, 64, 320))),
    ("perfauto_conv_64_320", case_perf, dict(M=0, N=320, K=0, conv=(4, 64, 64, 320), cta_pair=2)),
    ("perfauto_qkv_320", case_perf, dict(M=16384, N=320, K=320, cta_pair=2, graph=True)),
    ("perf_conv_16_1280", case_perf, dict(M=0, N=1280, K=0, conv=(4, 16, 16, 1280))),
]


def main():
    if len(sys.argv) >= 3 and sys.argv[1] == "--case":
        name = sys.argv[2]
        for n, fn, kw in CASES:
            if n == name:
                res = fn(**kw)
                print("RESULT " + json.dumps(res))
                return
        raise SystemExit(f"unknown case {name}")
    only = sys.argv[1:] if len(sys.argv) > 1 else None
    out_dir = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out_dir, exist_ok=True)
    results = {}
    for name, _, kw in CASES:
        if only and not any(name.startswith(o) for o in only):
            continue
        t0 = time.time()
        try:
            pr = subprocess.run([sys.executable, os.path.abspath(__file__), "--case", name],
                                capture_output=True, text=True, timeout=60)
            line = [l for l in pr.stdout.splitlines() if l.startswith("RESULT ")]
            if pr.returncode == 0 and line:
                res = json.loads(line[-1][7:])
            else:
                res = {"ok": False, "rc": pr.returncode, "stderr": pr.stderr[-1500:], "stdout": pr.stdout[-500:]}
        except subprocess.TimeoutExpired:
            res = {"ok": False, "timeout": True}
        res["secs"] = round(time.time() - t0, 1)
        res["args"] = {k: v for k, v in kw.items()}
        results[name] = res
        print(name, json.dumps(res)[:300], flush=True)
        with open(os.path.join(out_dir, "gemm_cases.json"), "w") as f:
            json.dump(results, f, indent=1)
    nfail = sum(1 for r in results.values() if not r.get("ok"))
    print(f"SUMMARY {len(results) - nfail}/{len(results)} ok")


if __name__ == "__main__":
    main()
