"""Data-parallel parity on real GPUs: R-GPU iterations at global batch B vs 1-GPU iterations at
batch B on the same seeds.  Checked: (1) every rank holds bit-identical LoRA weights after the steps (the DP
invariant), (2) the loss sequence agrees within 5 % + the bf16 loss floor (the two runs use different per-launch
batch sizes, hence different tile shapes and bf16 rounding), (3) the accumulated AdamW update points the same
way (cosine >= 0.9; element-wise comparison is meaningless because Adam turns a sign flip of a near-zero
gradient into a full +-lr step).  Launch:

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
      tests/gpu_checks/dp_check.py
"""
import contextlib
import io
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def build(arch, dev, rank, world, batch, graphs):
    import torch
    from leco_b200.lora import LoRANetwork
    from leco_b200.scheduler import DDIMScheduler
    from leco_b200.synthetic import build_engine, prompt_embedding
    from leco_b200.trainer import LecoTrainer, PromptPair
    from leco_b200.unet import SPECS
    unet = build_engine(arch, dev, seed=0)
    torch.manual_seed(1234)
    with contextlib.redirect_stdout(io.StringIO()):
        net = LoRANetwork(unet, rank=4, multiplier=1.0, alpha=1.0)
    net.to(dev, dtype=torch.bfloat16)
    D = SPECS[arch].cross_attention_dim
    emb = {p: prompt_embedding(p, D) for p in ("van gogh", "", "painting")}
    pair = PromptPair(target=emb["van gogh"], positive=emb["van gogh"], unconditional=emb[""], neutral=emb["painting"],
                      guidance_scale=1.0, resolution=128, batch_size=batch, action="erase")
    tr = LecoTrainer(unet, net, DDIMScheduler("v_prediction"), [pair], lr=1e-3, max_denoising_steps=8, device=dev,
                     rank=rank, world_size=world, use_cuda_graphs=graphs, state_fp32=True)
    return tr, net


def main():
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    arch, batch, iters = "tiny21", 4, 3
    tr, net = build(arch, dev, rank, world, batch, graphs=True)
    flat0 = net.flat.params.float().clone()
    torch.manual_seed(7)
    losses = [tr.iteration().item() for _ in range(iters)]
    flat = net.flat.params.float().clone()
    # all ranks must hold identical weights after identical updates
    ref = flat.clone()
    dist.broadcast(ref, 0)
    same = bool(torch.equal(ref, flat))
    ok_all = torch.tensor([int(same)], device=dev)
    dist.all_reduce(ok_all, op=dist.ReduceOp.MIN)
    result = None
    if rank == 0:
        tr1, net1 = build(arch, dev, 0, 1, batch, graphs=False)
        torch.manual_seed(7)
        losses1 = [tr1.iteration().item() for _ in range(iters)]
        flat1 = net1.flat.params.float()
        mask = net1.flat.mask.bool()
        # two bf16 executions with different per-launch batch sizes (tile shapes, rounding order): the same loss up to
        # 5 % plus the rounding variance floor of the objective measured in round 1 (1e-4, profiles/r1_dp_check_w2.json)
        BF16_LOSS_FLOOR = 1e-4
        up_dp, up_1 = (flat - flat0)[mask], (flat1 - flat0)[mask]
        cos = float(torch.dot(up_dp, up_1) / (up_dp.norm() * up_1.norm() + 1e-30))
        result = {"world": world, "losses_dp": losses, "losses_single": losses1,
                  "loss_rel_err": max(abs(a - b) / abs(b) for a, b in zip(losses, losses1)),
                  "loss_ok": all(abs(a - b) <= 0.05 * abs(b) + BF16_LOSS_FLOOR for a, b in zip(losses, losses1)),
                  "update_cosine": cos, "update_max_abs_diff": float((up_dp - up_1).abs().max()),
                  "ranks_identical": bool(ok_all.item())}
        result["ok"] = result["loss_ok"] and result["ranks_identical"] and cos >= 0.9
        print("DP_RESULT " + json.dumps(result))
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        json.dump(result, open(os.path.join(ROOT, "gpurun_out", f"dp_check_w{world}.json"), "w"), indent=1)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
