"""bench.py — LECO training-step throughput on B200 (metric of BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config NAME] [--k 25]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one LECO iteration (train_lora.py:141-302): k DDIM denoise steps under CFG with LoRA on, the
positive/neutral/unconditional/target predictions, the erase MSE, backward into the LoRA matrices and AdamW.
Workloads (`--config`, synthetic seeded weights / prompt embeddings — no checkpoints exist offline):
  sd21        BASELINE configs[1]: SD2.1 UNet, lierla rank 4, bf16, 512 px, prompt batch 2 per GPU   (default at N=1)
  sd21_b4     BASELINE configs[4]: same with prompt batch 4 per GPU (global 32 on 8 GPUs)            (default at N>1)
  sd15_c3lier BASELINE configs[2]: SD1.5 UNet, c3lier rank 8 (attention + conv adapters), batch 4
  sdxl        BASELINE configs[3]: SDXL UNet through the XL loop, lierla rank 4, 1024 px, batch 2
k is fixed to 25 (= E[k] of the reference's uniform draw over [1,49]; SURVEY §8d asks for the fixed-k variant); the
line also carries `random_k`: the same measurement with the reference's own seeded draw k ~ U[1,49].
value = latents/s = global batch x iterations/s, inputs resident in HBM.  e2e = the same through
LecoTrainer.iteration() with host noise (pinned, H2D) and a D2H read of the loss every step.  Weak scaling.

--impl reference times the reference's CPU path (oracle port: LECO arithmetic pinned against the reference's train
loop + restated diffusers UNet, fp32, all host threads): COMPLETE iterations (LoRA on, backward, AdamW) at a
bounded k, each phase timed; the k of the workload is reached by scaling the measured per-denoise-step time only.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# name -> (arch, train method, rank, per-GPU batch, resolution, prediction type, F per-sample fwd TFLOP, F_bwd/F)
CONFIGS = {
    "sd21": dict(arch="sd21", method="lierla", rank=4, batch=2, res=512, pred="v_prediction", f_fwd=0.804, f_bwd=1.23,
                 baseline="configs[1]"),
    "sd21_b4": dict(arch="sd21", method="lierla", rank=4, batch=4, res=512, pred="v_prediction", f_fwd=0.804, f_bwd=1.23,
                    baseline="configs[4] (per-GPU share: global 32 on 8 GPUs)"),
    "sd15_c3lier": dict(arch="sd15", method="c3lier", rank=8, batch=4, res=512, pred="epsilon", f_fwd=0.804, f_bwd=1.23,
                        baseline="configs[2]"),
    "sdxl": dict(arch="sdxl", method="lierla", rank=4, batch=2, res=1024, pred="epsilon", f_fwd=6.761, f_bwd=1.17,
                 baseline="configs[3]"),
    # reduced-width twin for exercising this file's plumbing in seconds (tests/test_bench_contract_cpu.py); not a result
    "tiny21": dict(arch="tiny21", method="lierla", rank=4, batch=2, res=128, pred="v_prediction", f_fwd=0.002, f_bwd=1.23,
                   baseline="(plumbing check, reduced-width twin)"),
}


def default_workload(n_gpus: int) -> str:
    """BASELINE configs[1] on one GPU (the configuration the metric is quoted on), configs[4] (prompt batch 4 per GPU,
    global 32 on 8) when sharded — for BOTH arms, so the driver's ratio compares like with like."""
    return "sd21" if n_gpus == 1 else "sd21_b4"


def w_min_tflop(cfg, k: int, distinct_nograd: int = 2) -> float:
    """Non-redundant work per iteration: 2B*F*k + (distinct+1)*B*F + B*F_bwd."""
    b, f = cfg["batch"], cfg["f_fwd"]
    return (2 * b * k + (distinct_nograd + 1) * b) * f + b * f * cfg["f_bwd"]


def w_ref_tflop(cfg, k: int) -> float:
    """As executed by the reference: 2B*F*(k+4) + 2B*F_bwd."""
    b, f = cfg["batch"], cfg["f_fwd"]
    return 2 * b * (k + 4) * f + 2 * b * f * cfg["f_bwd"]


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("bf16_tflops", 1590.0), d.get("bf16_tflops_sustained", 1400.0), "measured"
    return 1590.0, 1400.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-i", str(index), "-lms", "200"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [l.strip().split(",") for l in open(self.f.name) if l.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for nm, v in zip(names, r[2:6]):
                    if v.strip().lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def config_dict(name, cfg, n, k, b_global):
    method = "lierla" if cfg["method"] == "lierla" else "c3lier (attention + conv adapters)"
    return {"workload": f"BASELINE {cfg['baseline']}: {cfg['arch'].upper()}-arch UNet (synthetic weights), {method} rank "
                        f"{cfg['rank']}, bf16, {cfg['res']}px, prompt batch {cfg['batch']}/GPU, {cfg['pred']} DDIM, "
                        f"max_denoising_steps=50, k fixed {k}",
            "name": name, "global_batch": b_global, "k_denoise": k, "parallelism": f"dp{n}",
            "l2": "per-iteration working set (>= 1.7 GB weights + activations) >> 126 MB L2; no flush needed"}


# ----------------------------------------------------------------------------- reference arm (CPU, oracle port)
def _pick_threads(cores: int) -> int:
    """Host thread count for the CPU arm: all cores unless a quick conv probe says oversubscription hurts
    (it does on 100+-thread boxes)."""
    import torch
    best, best_t = cores, None
    x = torch.randn(2, 320, 64, 64)
    w = torch.randn(320, 320, 3, 3)
    for n in sorted({cores, max(1, cores // 2), max(1, cores // 4)}, reverse=True):
        torch.set_num_threads(n)
        torch.nn.functional.conv2d(x, w, padding=1)
        t0 = time.perf_counter()
        for _ in range(3):
            torch.nn.functional.conv2d(x, w, padding=1)
        dt = time.perf_counter() - t0
        if best_t is None or dt < 0.9 * best_t:
            best, best_t = n, dt
    return best


def cpu_reference_iterations(cfg, n_iter: int, k_meas: int, batch: int):
    """COMPLETE LECO iterations of the reference's CPU path at this architecture / resolution: oracle/leco_ref.leco_iteration
    (pinned bit-for-bit against the reference's unmodified train_lora.train(), tests/test_oracle_pinned.py) driving the
    restated diffusers UNet in fp32 on all host threads — k_meas denoise steps under CFG with LoRA on, the three LoRA-off
    predictions, the LoRA-on prediction with autograd, loss, backward, torch.optim.AdamW.  Returns per-iteration wall
    times of the denoise loop and of the whole iteration."""
    import contextlib
    import io
    import torch
    from leco_b200.synthetic import prompt_embedding
    from oracle import leco_ref
    from oracle.sched_ref import create_noise_scheduler
    from oracle.unet_ref import CONFIGS as UCFG, build_unet
    if cfg["arch"] == "sdxl":
        raise SystemExit("the CPU arm covers the SD1.x/2.x loop (train_lora.py); use --config sd21 / sd15_c3lier")
    cores = _pick_threads(os.cpu_count() or 1)
    torch.set_num_threads(cores)
    unet = build_unet(cfg["arch"])
    D = UCFG[cfg["arch"]].cross_attention_dim
    emb = {p: prompt_embedding(p, D) for p in ("van gogh", "")}
    torch.manual_seed(1234)
    targets = leco_ref.ATTN_TARGETS + (leco_ref.CONV_TARGETS if cfg["method"] == "c3lier" else [])
    with contextlib.redirect_stdout(io.StringIO()):
        net = leco_ref.LoRANetworkRef(unet, rank=cfg["rank"], multiplier=1.0, alpha=1.0, targets=targets)
    pair = leco_ref.PromptPairRef(target=emb["van gogh"], positive=emb["van gogh"], unconditional=emb[""],
                                  neutral=emb[""], guidance_scale=1.0, resolution=cfg["res"], batch_size=batch, action="erase")
    opt = torch.optim.AdamW(net.prepare_optimizer_params(), lr=1e-4)
    lrs = torch.optim.lr_scheduler.ConstantLR(opt, factor=1)
    sched = create_noise_scheduler("ddim", cfg["pred"])
    t_den, t_tot = [], []
    for _ in range(n_iter):
        rec = {"clock": time.perf_counter}
        leco_ref.leco_iteration(unet, sched, net, opt, lrs, [pair], max_denoising_steps=50, fixed_k=k_meas, record=rec)
        t_den.append(rec["t_denoise"])
        t_tot.append(rec["t_total"])
    return t_den, t_tot, cores


def cpu_projection(cfg, t_den, t_tot, k_meas, batch_meas, k):
    """latents/s of the CPU path on the workload's (batch, k) from measured phases: iteration(k) = k * (measured time of
    one denoise step) + (measured rest of the iteration); the CPU path does not batch-amortise, so time scales with the
    prompt batch (latents/s is batch-invariant)."""
    step_s = sum(t_den) / len(t_den) / k_meas
    tail_s = sum(t_tot) / len(t_tot) - step_s * k_meas
    t_iter_meas = sum(t_tot) / len(t_tot)
    t_iter_k = k * step_s + tail_s
    return {"measured": {"k": k_meas, "prompt_batch": batch_meas, "iterations": len(t_tot), "s_per_iteration": t_iter_meas,
                         "s_denoise_step": step_s, "s_rest_of_iteration": tail_s,
                         "latents_per_s": batch_meas / t_iter_meas},
            "projected": {"k": k, "s_per_iteration_at_measured_batch": t_iter_k, "latents_per_s": batch_meas / t_iter_k,
                          "how": "k x measured denoise-step time + measured rest (3 LoRA-off + 1 LoRA-on forward, backward, AdamW)"}}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    name = args.config or default_workload(args.gpus)     # the same default workload as our arm at this N
    cfg = CONFIGS[name]
    k = args.k
    if cfg["arch"] == "sdxl":
        # one complete fp32 iteration of the XL loop at 1024 px is ~48 TFLOP on the host cores (tens of minutes): no
        # bounded sample of it fits the run; the driver's reference runs use the default workloads (sd21 / sd21_b4)
        print(json.dumps({"impl": "reference", "unavailable": "the CPU arm covers the SD1.x/2.x loop (train_lora.py): one fp32 "
                          "SDXL iteration at 1024 px does not fit a bounded sample; use --config sd21 / sd21_b4 / sd15_c3lier"}))
        return
    n_iter = 2 if args.steps >= 2 else 1    # complete iterations at k=1: minutes, not hours
    # bounded sample of the workload: prompt batch <= 2 per measured iteration (the CPU path does not batch-amortise:
    # its time is linear in the batch, so latents/s is batch-invariant; cpu_projection states the batch it measured)
    b_meas = min(cfg["batch"], 2)
    t_den, t_tot, cores = cpu_reference_iterations(cfg, n_iter, 1, b_meas)
    pr = cpu_projection(cfg, t_den, t_tot, 1, b_meas, k)
    val = pr["projected"]["latents_per_s"]
    b_global = cfg["batch"] * args.gpus
    sample = (f"{n_iter} COMPLETE iteration(s) of the oracle port at k=1, prompt batch {b_meas} (CFG batch {2 * b_meas}): "
              f"denoise step + 3 LoRA-off + 1 LoRA-on forwards, loss, backward, AdamW = {pr['measured']['s_per_iteration']:.1f} s each "
              f"on {cores} host threads; value = workload k={k} from the measured phases (k x {pr['measured']['s_denoise_step']:.1f} s "
              f"+ {pr['measured']['s_rest_of_iteration']:.1f} s)")
    line = {"impl": "reference", "metric": "leco_train_latents_per_sec", "value": val, "unit": "latents/s",
            "value_kind": "projected to the workload's k from measured phases of complete iterations (see cpu_baseline)",
            "n_gpus": args.gpus, "steps": n_iter, "steps_requested": args.steps, "warmup": 0,
            "ms_per_step": 1000.0 * cfg["batch"] / val, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "fp32", "data": "synthetic", "config": config_dict(name, cfg, args.gpus, k, b_global),
            "cpu_baseline": {"value": val, "unit": "latents/s", "cores": cores, "kind": "port", "sample": sample, **pr},
            "e2e": {"value": val, "unit": "latents/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ----------------------------------------------------------------------------- our arm
def build_trainer(name, cfg, dev, rank, world, batch_global):
    import contextlib
    import io
    import torch
    import leco_b200.lora as plora
    from leco_b200.lora import LoRANetwork
    from leco_b200.scheduler import DDIMScheduler
    from leco_b200.synthetic import build_engine, prompt_embedding
    from leco_b200.trainer import EmbedsXL, LecoTrainer, PromptPair
    from leco_b200.unet import SPECS
    unet = build_engine(cfg["arch"], dev, seed=0)
    torch.manual_seed(1234)  # identical on every rank: adapter init + the loop's CPU draws
    saved = list(plora.DEFAULT_TARGET_REPLACE)
    try:
        if cfg["method"] == "c3lier":     # train_lora.py:44-46 (the reference extends the list in place)
            plora.DEFAULT_TARGET_REPLACE += plora.UNET_TARGET_REPLACE_MODULE_CONV
        with contextlib.redirect_stdout(io.StringIO()):
            net = LoRANetwork(unet, rank=cfg["rank"], multiplier=1.0, alpha=1.0, train_method="full")
    finally:
        plora.DEFAULT_TARGET_REPLACE[:] = saved
    net.to(dev, dtype=torch.bfloat16)
    spec = SPECS[cfg["arch"]]
    D = spec.cross_attention_dim
    emb = {p: prompt_embedding(p, D) for p in ("van gogh", "")}
    if spec.text_time:
        emb = {p: EmbedsXL(e, prompt_embedding(p + "/pooled", spec.add_text_dim)[0, :1]) for p, e in emb.items()}
    pair = PromptPair(target=emb["van gogh"], positive=emb["van gogh"], unconditional=emb[""], neutral=emb[""],
                      guidance_scale=1.0, resolution=cfg["res"], batch_size=batch_global, action="erase")  # examples/prompts.yaml
    trainer = LecoTrainer(unet, net, DDIMScheduler(cfg["pred"]), [pair], lr=1e-4, max_denoising_steps=50, device=dev,
                          rank=rank, world_size=world)
    return trainer, net


def _event_time(fn, reps, flush=None):
    """mean device ms of fn() over `reps` runs (CUDA events on the launching stream, optional L2 flush between runs)."""
    import torch
    ts = []
    for i in range(reps + 3):
        if flush is not None:
            flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        if i >= 3:
            ts.append(e0.elapsed_time(e1))
    return sum(ts) / len(ts)


def kernel_rooflines(cfg, dev, burst, how):
    """Live measurements (CUDA events, L2 flushed between launches) of the kernels that dominate a denoise forward:
    the fused flash-attention kernel at the top level's self-attention shape (the single largest kernel by time in the
    ncu launch list, profiles/), the tcgen05 implicit-GEMM conv, and the whole attention path (projections + cores,
    self and cross, all levels of one forward) = the 'attention-GEMM' figure north_star asks for."""
    import torch
    from leco_b200 import ops
    from leco_b200.unet import SPECS
    spec = SPECS[cfg["arch"]]
    n_s = 2 * cfg["batch"]
    lat = cfg["res"] // 8
    flush = torch.empty(256 << 20, device=dev, dtype=torch.uint8)   # 256 MiB > 126 MB L2
    out = {}

    def traffic(fname):
        p = os.path.join(ROOT, "profiles", fname)
        return json.load(open(p)).get("traffic_bytes_per_launch") if os.path.exists(p) else None

    # ---- transformer blocks per resolution level: (tokens per sample, channels, heads, depth, #Transformer2DModel)
    # down blocks hold layers_per_block transformers, up blocks layers_per_block + 1, the mid block one (at the lowest
    # resolution with the last level's width)
    L = len(spec.block_out_channels)
    levels = []
    hw = lat * lat
    for i, c in enumerate(spec.block_out_channels):
        n_tr = (2 * spec.layers_per_block + 1) if spec.attn_levels[i] else 0
        if i == L - 1:
            n_tr += 1
        if n_tr:
            levels.append((hw, c, spec.num_heads[i], spec.transformer_depth[i], n_tr))
        if i != L - 1:
            hw //= 4
    # ---- flash attention at the largest self-attention shape
    S, C, H, _, _ = levels[0]
    d = C // H
    if d <= 64:
        qkv = torch.randn((n_s * S, 3 * C), device=dev).to(torch.bfloat16)
        ms = _event_time(lambda: ops.flash_attention(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], n_s, S, S, H, d, d ** -0.5),
                         6, flush)
        fl = 4.0 * n_s * H * S * S * d
        tf = fl / (ms * 1e-3) / 1e12
        out["roofline"] = {"bound": "tensor", "kernel": f"flash_attn_fwd_ts_kernel softmax(QK^T/sqrt d)V, S={S}, d={d}, {n_s * H} (sample, head) "
                           f"pairs — the time-dominant kernel of a denoise forward (profiles/ launch list)",
                           "achieved": tf, "peak": burst, "unit": "TFLOP/s", "frac": tf / burst,
                           "traffic": traffic("flash_kernel_traffic.json"), "peak_source": how, "ms": ms,
                           "algorithmic_flop_per_launch": fl}
    # ---- implicit-GEMM conv at the top level
    c0 = spec.block_out_channels[0]
    xa = torch.randn((n_s * lat * lat, c0), device=dev).to(torch.bfloat16)
    wk = (torch.randn((c0, 9 * c0), device=dev) * 0.02).to(torch.bfloat16)
    yo = torch.empty((n_s * lat * lat, c0), device=dev, dtype=torch.bfloat16)
    ms = _event_time(lambda: ops.gemm(xa, wk, yo, conv_nhw=(n_s, lat, lat)), 6, flush)
    fl = 2.0 * n_s * lat * lat * c0 * 9 * c0
    tf = fl / (ms * 1e-3) / 1e12
    conv = {"bound": "tensor", "kernel": f"tcgen05 implicit-GEMM conv3x3 {c0}->{c0} @{lat}x{lat}, {n_s} samples "
            f"(M={n_s * lat * lat},N={c0},K={9 * c0})", "achieved": tf, "peak": burst, "unit": "TFLOP/s", "frac": tf / burst,
            "traffic": traffic("conv_kernel_traffic.json"), "peak_source": how, "ms": ms}
    if "roofline" in out:
        out["roofline_conv"] = conv
    else:
        out["roofline"] = conv
    # ---- attention path of one forward: per level, one transformer block's attention chain timed back to back
    # (L2-warm between the kernels of a chain as in the real forward; flushed between repetitions), weighted by the
    # number of blocks at that level
    tot_ms = tot_fl = 0.0
    Dctx = spec.cross_attention_dim
    for (S, C, H, depth, n_tr) in levels:
        d = C // H
        if d > 64:
            return out   # SD1.x head dims 80/160: attention core runs the materialised path (no fused-kernel figure)
        M = n_s * S
        x = torch.randn((M, C), device=dev).to(torch.bfloat16)
        ctx = torch.randn((n_s * 77, Dctx), device=dev).to(torch.bfloat16)
        wqkv = (torch.randn((3 * C, C), device=dev) * 0.03).to(torch.bfloat16)
        wo = (torch.randn((C, C), device=dev) * 0.03).to(torch.bfloat16)
        wkv = (torch.randn((2 * C, Dctx), device=dev) * 0.03).to(torch.bfloat16)
        bo = torch.zeros(C, device=dev, dtype=torch.bfloat16)

        def chain():
            qkv = ops.gemm(x, wqkv)
            a = ops.flash_attention(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], n_s, S, S, H, d, d ** -0.5)
            h1 = ops.gemm(a, wo, bias=bo, residual=x)
            q2 = ops.gemm(h1, wo)
            kv = ops.gemm(ctx, wkv)
            a2 = ops.flash_attention(q2, kv[:, :C], kv[:, C:], n_s, S, 77, H, d, d ** -0.5)
            ops.gemm(a2, wo, bias=bo, residual=h1)
        ms = _event_time(chain, 5, flush)
        fl = 2.0 * M * C * 6 * C + 2.0 * n_s * 77 * Dctx * 2 * C + 4.0 * n_s * H * S * (S + 77) * d
        tot_ms += ms * n_tr * depth
        tot_fl += fl * n_tr * depth
    tf = tot_fl / (tot_ms * 1e-3) / 1e12
    out["roofline_attention"] = {"bound": "tensor", "scope": "attention path of ONE denoise forward: q/k/v/out projections + fused "
                                 "softmax(QK^T)V cores, self and cross, every transformer block (frozen weights; LoRA K-segment excluded)",
                                 "achieved": tf, "peak": burst, "unit": "TFLOP/s", "frac": tf / burst, "ms_per_forward": tot_ms,
                                 "tflop_per_forward": tot_fl / 1e12, "peak_source": how}
    return out


def scaling_base(name, dev, k, steps):
    """One GPU on the per-GPU share of the sharded workload (`name`, = what every rank of an N > 1 run executes minus
    the all-reduce): latents/s with the latent noise resident in HBM, CUDA events, 3 warm-up iterations."""
    import torch
    cfg = CONFIGS[name]
    lat = cfg["res"] // 8
    tr, _ = build_trainer(name, cfg, dev, 0, 1, cfg["batch"])
    for _ in range(3):
        tr.iteration(fixed_k=k)
    noise = torch.randn((cfg["batch"], 4, lat, lat), device=dev)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        tr.iteration(fixed_k=k, device_noise=noise)
    e1.record()
    torch.cuda.synchronize()
    sec = e0.elapsed_time(e1) / 1000.0
    return {"workload": config_dict(name, cfg, 1, k, cfg["batch"])["workload"], "name": name, "n_gpus": 1,
            "value": cfg["batch"] * steps / sec, "unit": "latents/s", "steps": steps, "ms_per_step": 1000.0 * sec / steps,
            "note": "N > 1 lines run this workload per GPU: scaling efficiency at N = value(N) / (N x this value)"}


def run_ours(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torchrun for --gpus > 1")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    name = args.config or default_workload(world)
    cfg = CONFIGS[name]
    k = args.k
    b_local = cfg["batch"]
    b_global = b_local * world
    lat = cfg["res"] // 8
    trainer, net = build_trainer(name, cfg, dev, rank, world, b_global)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(tr, n_iter, e2e: bool, fixed_k=k, dp=True):
        """returns (seconds = max over ranks of device time, last loss)."""
        noise = torch.randn((b_local, 4, lat, lat), device=dev) if not e2e else None
        if dp:
            barrier()
        else:
            torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        loss_val = None
        ks = []
        for _ in range(n_iter):
            loss = tr.iteration(fixed_k=fixed_k, device_noise=noise)
            ks.append(tr.last["k"])
            if e2e:
                loss_val = loss.item()          # D2H read of the step's result, every step
        e1.record()
        if dp:
            barrier()
        else:
            torch.cuda.synchronize()
        sec = torch.tensor([e0.elapsed_time(e1) / 1000.0], device=dev)
        if world > 1 and dp:
            dist.all_reduce(sec, op=dist.ReduceOp.MAX)
        return sec.item(), (loss_val if e2e else loss.item()), ks

    warm = max(3, args.warmup)
    for _ in range(warm):
        trainer.iteration(fixed_k=k)
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    trainer.launches = 0
    trainer._ar_events = []                # the warm-up all-reduces include NCCL's lazy communicator set-up
    trainer.profile_phases = True
    sec, loss_a, _ = timed(trainer, args.steps, e2e=False)
    phases = trainer.phase_ms()
    trainer.profile_phases = False
    launches = trainer.launches
    sec_e2e, loss_b, _ = timed(trainer, args.steps, e2e=True)
    h2d_bytes = trainer.h2d_bytes          # host noise of one step (pinned staging -> device), counted by the trainer
    clocks = sampler.stop() if sampler else None
    ar_ms = trainer.allreduce_ms()
    peak_mem = torch.cuda.max_memory_allocated(dev)
    # random-k variant: the reference's own draw k ~ U[1,49] (train_lora.py:154-156), seeded identically on all ranks
    torch.manual_seed(4321)
    n_rk = max(4, min(args.steps, 12))
    sec_rk, _, ks_rk = timed(trainer, n_rk, e2e=False, fixed_k=None)

    burst, sustained, how = measured_peaks()
    extra = {}
    if rank == 0 and not args.no_kernel_rooflines:
        try:
            extra = kernel_rooflines(cfg, dev, burst, how)
        except Exception as e:  # the step numbers stand on their own
            extra = {"roofline": None, "roofline_error": repr(e)}
    if world > 1:
        barrier()

    if rank == 0:
        it_s = args.steps / sec
        val = b_global * it_s
        val_e2e = b_global * args.steps / sec_e2e
        wmin = w_min_tflop(cfg, k) * world
        line = {
            "metric": "leco_train_latents_per_sec", "value": val, "unit": "latents/s", "n_gpus": world,
            "steps": args.steps, "warmup": warm, "ms_per_step": 1000.0 * sec / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": config_dict(name, cfg, world, k, b_global),
            "iterations_per_sec": it_s, "loss": loss_b,
            "e2e": {"value": val_e2e, "unit": "latents/s", "h2d_bytes_per_step": h2d_bytes,
                    "d2h_bytes_per_step": 4},
            "gpu_launches": int(launches),
            "random_k": {"value": b_global * n_rk / sec_rk, "unit": "latents/s", "steps": n_rk, "ms_per_step": 1000.0 * sec_rk / n_rk,
                         "k_drawn": ks_rk, "mean_k": sum(ks_rk) / len(ks_rk), "seed": 4321,
                         "note": "k ~ U[1,49] from the reference's own draw; ms_per_step scales with mean_k"},
            "step_roofline": {"bound": "tensor", "w_min_tflop_per_step": wmin, "w_ref_tflop_per_step":
                              w_ref_tflop(cfg, k) * world, "achieved": wmin / (sec / args.steps),
                              "peak": sustained * world, "unit": "TFLOP/s",
                              "frac": wmin / (sec / args.steps) / (sustained * world), "peak_source": how},
            "phases": None if phases is None else {
                **phases, "denoise_step_ms": phases["denoise_loop_ms"] / max(1.0, phases["mean_k"]),
                "what": "CUDA events inside the timed region: k-step CFG denoise loop (LoRA on) | 3 LoRA-off + 1 LoRA-on "
                        "predictions, loss, backward, optimizer"},
            "peak_memory_bytes": int(peak_mem),
            "clocks": clocks,
        }
        line.update(extra)
        if world > 1:
            # event-to-event on rank 0: the collective itself plus the wait for the slowest rank to arrive
            line["allreduce"] = {"device_us_per_step": None if ar_ms is None else 1000.0 * ar_ms,
                                 "bytes": int(net.flat.grads_ext.numel() * 4),
                                 "what": "ONE ncclAllReduce of the flat fp32 LoRA gradient with the loss in its last slot"}
        if world == 1 and args.config is None and not args.no_scaling_base:
            # The sharded runs (N > 1) use configs[4]'s per-GPU share (prompt batch 4 per GPU), this N=1 line configs[1]
            # (prompt batch 2): the like-for-like base of the 1 -> N curve is the configs[4] share on ONE GPU, measured
            # here after everything else (a failure cannot touch the numbers above).
            try:
                line["scaling_base"] = scaling_base(default_workload(2), dev, k, max(3, min(args.steps, 5)))
            except Exception as e:
                line["scaling_base"] = {"error": repr(e)}
        if world == 1 and not args.no_cpu_baseline and cfg["arch"] != "sdxl":
            # bounded sample: ONE complete iteration of the CPU path at k=1, prompt batch 1
            t_den, t_tot, cores = cpu_reference_iterations(cfg, 1, 1, 1)
            pr = cpu_projection(cfg, t_den, t_tot, 1, 1, k)
            line["cpu_baseline"] = {"value": pr["projected"]["latents_per_s"], "unit": "latents/s", "cores": cores, "kind": "port",
                                    "sample": f"1 COMPLETE iteration of the oracle port (k=1, prompt batch 1: denoise step, 3 LoRA-off + "
                                              f"1 LoRA-on forwards, loss, backward, AdamW) = {pr['measured']['s_per_iteration']:.1f} s on {cores} "
                                              f"host threads; value = k={k} from the measured phases", **pr}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default=None, choices=sorted(CONFIGS), help="workload (default: sd21 at N=1, sd21_b4 at N>1)")
    ap.add_argument("--k", type=int, default=25, help="fixed number of denoise steps per iteration")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-kernel-rooflines", action="store_true")
    ap.add_argument("--no-scaling-base", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
