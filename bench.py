"""bench.py — LECO training-step throughput on B200 (metric of BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one LECO iteration (train_lora.py:141-302): k DDIM denoise steps under CFG with
LoRA on, the positive/neutral/unconditional/target predictions, the erase MSE, backward into
the LoRA matrices and AdamW.  Workload (config.workload): BASELINE configs[1] = SD2.1 UNet
architecture, lierla rank 4, bf16, 512 px (64x64 latents), prompt batch 2 per GPU, k fixed to
25 (= E[k] of the reference's uniform draw over [1,49]; SURVEY §8d asks for the fixed-k
variant), synthetic seeded weights / prompt embeddings (no checkpoints exist offline).
value = latents/s = global batch x iterations/s.  Weak scaling: per-GPU batch fixed.

--impl reference times the reference's CPU path (oracle port: LECO arithmetic + restated
diffusers UNet, fp32, all host threads) on a bounded sample of the same workload.
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

ARCH = "sd21"
RES = 512
B_PER_GPU = 2
RANK_LORA = 4
F_FWD_TFLOP = 0.804        # per-sample UNet forward, SD2.1 @64x64 latent
F_BWD_FACTOR = 1.23        # backward-data ~ 1.23 F


def w_min_tflop(b: int, k: int, distinct_nograd: int = 2) -> float:
    """Non-redundant work per iteration: 2B*F*k + (distinct+1)*B*F + B*F_bwd."""
    return (2 * b * k + (distinct_nograd + 1) * b) * F_FWD_TFLOP + b * F_FWD_TFLOP * F_BWD_FACTOR


def w_ref_tflop(b: int, k: int) -> float:
    """As executed by the reference: 2B*F*(k+4) + 2B*F_bwd."""
    return 2 * b * (k + 4) * F_FWD_TFLOP + 2 * b * F_FWD_TFLOP * F_BWD_FACTOR


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


# ----------------------------------------------------------------------------- reference arm
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


def cpu_reference_forward_seconds(n_fwd: int, warm: int, b: int = 1):
    """Times `predict_noise` of the oracle port (LECO arithmetic + restated UNet, fp32, host threads) for a CFG
    batch of 2*b samples at the bench architecture/resolution.  n_fwd is capped so a run stays within minutes."""
    import torch
    from oracle import leco_ref
    from oracle.sched_ref import create_noise_scheduler
    from oracle.unet_ref import CONFIGS, build_unet
    cores = _pick_threads(os.cpu_count() or 1)
    torch.set_num_threads(cores)
    unet = build_unet(ARCH)
    sched = create_noise_scheduler("ddim", "v_prediction")
    sched.set_timesteps(50)
    g = torch.Generator().manual_seed(0)
    lat = torch.randn((b, 4, RES // 8, RES // 8), generator=g)
    emb = torch.randn((2 * b, 77, CONFIGS[ARCH].cross_attention_dim), generator=g)
    times = []
    with torch.no_grad():
        for i in range(warm + n_fwd):
            t0 = time.perf_counter()
            leco_ref.predict_noise(unet, sched, sched.timesteps[0], lat, emb, guidance_scale=3)
            if i >= warm:
                times.append(time.perf_counter() - t0)
    return sum(times) / len(times), cores


def cpu_latents_per_s(sec_per_cfg_fwd_b1: float, b: int, k: int) -> float:
    """Extrapolate one full reference iteration from the measured CFG forward (2 samples):
    (k+4) CFG forwards at batch b + backward (2b samples x 1.23 F)."""
    per_sample_fwd = sec_per_cfg_fwd_b1 / 2.0
    t_iter = 2 * b * (k + 4) * per_sample_fwd + 2 * b * F_BWD_FACTOR * per_sample_fwd
    return b / t_iter


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    k = args.k
    n_timed = max(1, min(args.steps, 2))  # one CFG forward is ~1 min of CPU: bound the run to a few minutes
    sec, cores = cpu_reference_forward_seconds(n_timed, 0)
    b_global = B_PER_GPU * args.gpus
    val = cpu_latents_per_s(sec, B_PER_GPU, k)  # CPU path does not shard: whole-job value on the host
    sample = (f"{n_timed} timed predict_noise call(s) (1 CFG UNet forward, 2 samples, fp32) of the same "
              f"arch/resolution = {sec:.1f} s each; iteration extrapolated as 2B(k+4) fwd + 2B*1.23 fwd, B={B_PER_GPU}, k={k}")
    line = {"impl": "reference", "metric": "leco_train_latents_per_sec", "value": val, "unit": "latents/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * B_PER_GPU / val,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
            "config": config_dict(args.gpus, k, b_global),
            "cpu_baseline": {"value": val, "unit": "latents/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "latents/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def config_dict(n, k, b_global):
    return {"workload": f"BASELINE configs[1]: SD2.1-arch UNet (synthetic weights), lierla rank {RANK_LORA}, bf16, "
                        f"{RES}px, prompt batch {B_PER_GPU}/GPU, v-pred DDIM, max_denoising_steps=50, k fixed {k}",
            "global_batch": b_global, "k_denoise": k, "parallelism": f"dp{n}",
            "l2": "per-iteration working set (1.7 GB weights + activations) >> 126 MB L2; no flush needed"}


# ----------------------------------------------------------------------------- our arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    from leco_b200 import capi
    from leco_b200.lora import LoRANetwork
    from leco_b200.scheduler import DDIMScheduler
    from leco_b200.synthetic import build_engine, prompt_embedding
    from leco_b200.trainer import LecoTrainer, PromptPair
    from leco_b200.unet import SPECS
    import contextlib
    import io

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

    k = args.k
    b_global = B_PER_GPU * world
    unet = build_engine(ARCH, dev, seed=0)
    torch.manual_seed(1234)  # identical on every rank: adapter init + the loop's CPU draws
    with contextlib.redirect_stdout(io.StringIO()):
        net = LoRANetwork(unet, rank=RANK_LORA, multiplier=1.0, alpha=1.0, train_method="full")
    net.to(dev, dtype=torch.bfloat16)
    D = SPECS[ARCH].cross_attention_dim
    emb = {p: prompt_embedding(p, D) for p in ("van gogh", "")}
    pair = PromptPair(target=emb["van gogh"], positive=emb["van gogh"], unconditional=emb[""], neutral=emb[""],
                      guidance_scale=1.0, resolution=RES, batch_size=b_global, action="erase")  # examples/prompts.yaml
    trainer = LecoTrainer(unet, net, DDIMScheduler("v_prediction"), [pair], lr=1e-4, max_denoising_steps=50,
                          device=dev, rank=rank, world_size=world)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(n_iter, e2e: bool):
        """returns (seconds = max over ranks of device time, last loss)."""
        noise = torch.randn((B_PER_GPU, 4, RES // 8, RES // 8), device=dev) if not e2e else None
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        loss_val = None
        for _ in range(n_iter):
            loss = trainer.iteration(fixed_k=k, device_noise=noise)
            if e2e:
                loss_val = loss.item()          # D2H read of the step's result, every step
        e1.record()
        barrier()
        sec = torch.tensor([e0.elapsed_time(e1) / 1000.0], device=dev)
        if world > 1:
            dist.all_reduce(sec, op=dist.ReduceOp.MAX)
        return sec.item(), (loss_val if e2e else loss.item())

    for _ in range(max(3, args.warmup)):
        trainer.iteration(fixed_k=k)
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    trainer.launches = 0
    sec, loss_a = timed(args.steps, e2e=False)
    launches = trainer.launches
    sec_e2e, loss_b = timed(args.steps, e2e=True)
    clocks = sampler.stop() if sampler else None

    # ---- roofline of the dominant kernel: the tcgen05 implicit-GEMM 3x3 conv at the 64x64 level
    from leco_b200 import ops
    n_s = 2 * B_PER_GPU
    xa = torch.randn((n_s * 64 * 64, 320), device=dev).to(torch.bfloat16)
    wk = (torch.randn((320, 9 * 320), device=dev) * 0.02).to(torch.bfloat16)
    out = torch.empty((n_s * 64 * 64, 320), device=dev, dtype=torch.bfloat16)
    flush = torch.empty(256 << 20, device=dev, dtype=torch.uint8)
    ts = []
    for i in range(8):
        flush.zero_()                                   # evict L2 (256 MiB > 126 MB) between launches
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.gemm(xa, wk, out, conv_nhw=(n_s, 64, 64))
        e1.record()
        torch.cuda.synchronize()
        if i >= 3:
            ts.append(e0.elapsed_time(e1))
    kern_ms = sum(ts) / len(ts)
    kern_flop = 2.0 * n_s * 64 * 64 * 320 * 9 * 320
    burst, sustained, how = measured_peaks()
    kern_tf = kern_flop / (kern_ms * 1e-3) / 1e12
    # DRAM traffic of one launch of this kernel from the committed `ncu --set full` capture (bytes, or None)
    traffic = None
    tpath = os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "conv_kernel_traffic.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get("traffic_bytes_per_launch")

    if rank == 0:
        it_s = args.steps / sec
        val = b_global * it_s
        val_e2e = b_global * args.steps / sec_e2e
        wmin = w_min_tflop(B_PER_GPU, k) * world
        line = {
            "metric": "leco_train_latents_per_sec", "value": val, "unit": "latents/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": 1000.0 * sec / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": config_dict(world, k, b_global),
            "iterations_per_sec": it_s, "loss": loss_b,
            "e2e": {"value": val_e2e, "unit": "latents/s", "h2d_bytes_per_step": trainer.h2d_bytes,
                    "d2h_bytes_per_step": 4},
            "gpu_launches": int(launches),
            "roofline": {"bound": "tensor", "kernel": "gemm_tcgen05_2cta_kernel<160> (library-chosen) implicit-GEMM conv3x3 320->320 @64x64, "
                         f"{n_s} samples (M=16384,N=320,K=2880)", "achieved": kern_tf, "peak": burst,
                         "unit": "TFLOP/s", "frac": kern_tf / burst, "traffic": traffic, "peak_source": how,
                         "ms": kern_ms},
            "step_roofline": {"bound": "tensor", "w_min_tflop_per_step": wmin, "w_ref_tflop_per_step":
                              w_ref_tflop(B_PER_GPU, k) * world, "achieved": wmin / (sec / args.steps),
                              "peak": sustained * world, "unit": "TFLOP/s",
                              "frac": wmin / (sec / args.steps) / (sustained * world), "peak_source": how},
            "clocks": clocks,
        }
        if world == 1 and not args.no_cpu_baseline:
            sec_f, cores = cpu_reference_forward_seconds(1, 0)
            line["cpu_baseline"] = {"value": cpu_latents_per_s(sec_f, B_PER_GPU, k), "unit": "latents/s", "cores": cores,
                                    "kind": "port",
                                    "sample": f"1 predict_noise call (CFG UNet fwd, 2 samples, fp32) = {sec_f:.1f} s on "
                                              f"{cores} host threads; iteration extrapolated as 2B(k+4)+2B*1.23 sample-forwards"}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--k", type=int, default=25, help="fixed number of denoise steps per iteration")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
