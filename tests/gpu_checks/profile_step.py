"""One LECO iteration of the bench workload inside a cudaProfilerStart/Stop window (for ncu
--profile-from-start off).  k is small (default 2): every denoise step is a replay of the same graph,
so the per-step launch list scales linearly in k.

  ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
      --log-file gpurun_out/launches.csv python tests/gpu_checks/profile_step.py --k 2
"""
import argparse
import contextlib
import io
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--k", type=int, default=2)
    ap.add_argument("--arch", default="sd21")
    ap.add_argument("--res", type=int, default=512)
    ap.add_argument("--batch", type=int, default=2)
    ap.add_argument("--no-graphs", action="store_true")
    args = ap.parse_args()
    import torch
    from leco_b200.lora import LoRANetwork
    from leco_b200.scheduler import DDIMScheduler
    from leco_b200.synthetic import build_engine, prompt_embedding
    from leco_b200.trainer import LecoTrainer, PromptPair
    from leco_b200.unet import SPECS
    dev = torch.device("cuda", 0)
    unet = build_engine(args.arch, dev, seed=0)
    torch.manual_seed(1234)
    with contextlib.redirect_stdout(io.StringIO()):
        net = LoRANetwork(unet, rank=4, multiplier=1.0, alpha=1.0)
    net.to(dev, dtype=torch.bfloat16)
    D = SPECS[args.arch].cross_attention_dim
    emb = {p: prompt_embedding(p, D) for p in ("van gogh", "")}
    pair = PromptPair(target=emb["van gogh"], positive=emb["van gogh"], unconditional=emb[""], neutral=emb[""],
                      guidance_scale=1.0, resolution=args.res, batch_size=args.batch, action="erase")
    tr = LecoTrainer(unet, net, DDIMScheduler("v_prediction"), [pair], lr=1e-4, max_denoising_steps=50,
                     use_cuda_graphs=not args.no_graphs)
    for _ in range(2):
        tr.iteration(fixed_k=args.k)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    loss = tr.iteration(fixed_k=args.k)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    print("loss", loss.item(), "k", args.k)


if __name__ == "__main__":
    main()
