"""`train(config, prompts)` — the reference's driver (train_lora.py:34-321) on the fused trainer.

    python -m leco_b200.train_lora --config_file examples/config.yaml

Same sequence: load models -> LoRANetwork(rank, alpha, training_method) with the c3lier target aliasing (:44-46) ->
optimizer from `train.optimizer` / `optimizer_args` ("k=v k=v", ast.literal_eval, :80-89) -> LR schedule (:90-95) ->
encode each distinct prompt once (:106-132) -> `iterations` passes of the loop body (LecoTrainer.iteration) -> periodic
saves that skip i == 0 and i == iterations-1, final `<name>_last.safetensors` (:292-309; SURVEY Q13: the weights are
saved in `train.precision`, `save.precision` is ignored, no metadata is written).  `config` / `prompts` may be the
reference's own pydantic objects or the dataclasses of leco_b200.config_util."""
from __future__ import annotations

import argparse
import ast
from pathlib import Path
from typing import Callable, List, Optional

import torch

from . import config_util, lora, model_util
from .lora import LoRANetwork
from .trainer import LecoTrainer, PromptPair


DEVICE_CUDA = torch.device("cuda:0")      # train_lora.py:26 (the mirror's train() takes `device=` instead)
NUM_IMAGES_PER_PROMPT = 1                 # train_lora_xl.py:29


def flush():
    """train_lora.py:29-31.  The reference calls it every iteration; the fused trainer never needs to (its graphs
    share one memory pool and nothing is freed inside the loop)."""
    import gc
    torch.cuda.empty_cache()
    gc.collect()


def parse_optimizer_args(s: Optional[str]) -> dict:
    """train_lora.py:81-87."""
    out = {}
    if s is not None and len(s) > 0:
        for arg in s.split(" "):
            key, value = arg.split("=")
            out[key] = ast.literal_eval(value)
    return out


def share_seed(device) -> int:
    """The data-parallel rule needs every rank to make the SAME draws (adapter init, prompt pair, k, global noise,
    SURVEY §8e); the reference never seeds, so rank 0's seed is broadcast and installed everywhere (torch.manual_seed
    also seeds the CUDA generator that the stochastic schedulers' step noise comes from)."""
    import torch.distributed as dist
    seed = torch.tensor([torch.initial_seed() % (1 << 62)], dtype=torch.int64, device=device)
    dist.broadcast(seed, src=0)
    torch.manual_seed(int(seed.item()))
    return int(seed.item())


def _json(obj) -> str:
    """`.json()` of the reference's pydantic objects, or the same for the dataclass mirror."""
    if hasattr(obj, "json"):
        return obj.json()
    import dataclasses
    import json
    return json.dumps(dataclasses.asdict(obj) if dataclasses.is_dataclass(obj) else vars(obj), default=str)


class RunLogger:
    """The logging side of the reference's driver: `metadata` printed under `logging.verbose` (train_lora.py:38-49),
    `wandb.init(project="LECO_<save.name>", config=metadata)` and one `wandb.log({"loss", "iteration", "lr"})` per
    iteration under `logging.use_wandb` (:51-52, :274-277).  Only rank 0 logs; `needs_loss` tells the loop whether it has
    to read the loss back (a device synchronisation the plain loop avoids)."""

    def __init__(self, config, prompts, rank: int = 0, wandb_module=None):
        self.metadata = {"prompts": ",".join(_json(p) for p in prompts), "config": _json(config)}
        self.verbose = bool(config.logging.verbose) and rank == 0
        self.wandb = None
        if self.verbose:
            print(self.metadata)
        if config.logging.use_wandb and rank == 0:
            if wandb_module is None:
                import wandb as wandb_module
            self.wandb = wandb_module
            self.wandb.init(project=f"LECO_{config.save.name}", config=self.metadata)

    @property
    def needs_loss(self) -> bool:
        return self.verbose or self.wandb is not None

    def iteration(self, i: int, loss: float, lr: float, k: int):
        if self.wandb is not None:
            self.wandb.log({"loss": loss, "iteration": i, "lr": lr})
        if self.verbose:
            print(f"iteration {i}: loss*1k {loss * 1000:.4f} lr {lr:.3e} k {k}")

    def finish(self):
        if self.wandb is not None and hasattr(self.wandb, "finish"):
            self.wandb.finish()


def train(config, prompts, *, device="cuda", rank: int = 0, world_size: int = 1,
          on_iteration: Optional[Callable[[int, float], None]] = None, xl: bool = False) -> List[float]:
    if world_size > 1:
        share_seed(device)
    log = RunLogger(config, prompts, rank)
    save_path = Path(config.save.path)
    weight_dtype = config_util.parse_precision(config.train.precision)
    save_weight_dtype = config_util.parse_precision(config.train.precision)      # sic
    if weight_dtype == torch.float16:
        raise NotImplementedError("train.precision: float16 is not supported (the reference's README calls it unstable); "
                                  "use bfloat16 (every example config) or float32")
    # bfloat16: adapters, gradients' rounding and optimizer state in bf16 like the reference's bf16 run (train_lora.py:78-89).
    # float32 (the notebook's setting): fp32 master adapters + fp32 optimizer state, saved as fp32; the frozen UNet still
    # runs on the bf16 tensor cores with fp32 accumulation (leco_b200.lora module docstring).
    if xl:
        tokenizer, text_encoder, unet, scheduler = model_util.load_models_xl(
            config.pretrained_model.name_or_path, scheduler_name=config.train.noise_scheduler, device=device)
        encode = model_util.encode_prompts_xl
    else:
        tokenizer, text_encoder, unet, scheduler = model_util.load_models(
            config.pretrained_model.name_or_path, scheduler_name=config.train.noise_scheduler,
            v2=config.pretrained_model.v2, v_pred=config.pretrained_model.v_pred, device=device)
        encode = model_util.encode_prompts
    unet.enable_xformers_memory_efficient_attention()      # train_lora.py:68 (a no-op: attention is the fused kernel)
    unet.requires_grad_(False)
    unet.eval()

    saved_targets = list(lora.DEFAULT_TARGET_REPLACE)
    try:
        if config.network.type == "c3lier":                 # train_lora.py:44-46 extends the module-level list in place
            lora.DEFAULT_TARGET_REPLACE += lora.UNET_TARGET_REPLACE_MODULE_CONV
        network = LoRANetwork(unet, rank=config.network.rank, multiplier=1.0, alpha=config.network.alpha,
                              train_method=config.network.training_method).to(device, dtype=weight_dtype)
    finally:
        lora.DEFAULT_TARGET_REPLACE[:] = saved_targets

    print("Prompts")
    cache = {}
    pairs = []
    for s in prompts:
        print(s)
        for p in (s.target, s.positive, s.neutral, s.unconditional):
            if p not in cache:
                cache[p] = encode(tokenizer, text_encoder, [p])      # train_lora.py:118-124: once per distinct prompt
        pairs.append(PromptPair(target=cache[s.target], positive=cache[s.positive], unconditional=cache[s.unconditional],
                                neutral=cache[s.neutral], guidance_scale=s.guidance_scale, resolution=s.resolution,
                                dynamic_resolution=s.dynamic_resolution, batch_size=s.batch_size, action=s.action,
                                dynamic_crops=getattr(s, "dynamic_crops", False)))
    del tokenizer, text_encoder                # train_lora.py:134-137: the text side is freed before the loop

    trainer = LecoTrainer(unet, network, scheduler, pairs, lr=config.train.lr, optimizer=config.train.optimizer,
                          optimizer_kwargs=parse_optimizer_args(config.train.optimizer_args),
                          lr_scheduler=config.train.lr_scheduler, iterations=config.train.iterations,
                          max_denoising_steps=config.train.max_denoising_steps, device=device, rank=rank,
                          world_size=world_size)
    losses = []
    for i in range(config.train.iterations):
        lr_now = trainer.optimizer.lr        # the rate this iteration steps with (= lr_scheduler.get_last_lr()[0], :276)
        loss = trainer.iteration()
        if on_iteration is not None or log.needs_loss:
            v = float(loss.item())           # reading the loss synchronises; the plain loop never does
            losses.append(v)
            if on_iteration is not None:
                on_iteration(i, v)
            log.iteration(i, v, lr_now, trainer.last["k"])
        if i % config.save.per_steps == 0 and i != 0 and i != config.train.iterations - 1 and rank == 0:
            print("Saving...")
            save_path.mkdir(parents=True, exist_ok=True)
            network.save_weights(str(save_path / f"{config.save.name}_{i}steps.safetensors"), dtype=save_weight_dtype)
    if rank == 0:
        print("Saving...")
        save_path.mkdir(parents=True, exist_ok=True)
        network.save_weights(str(save_path / f"{config.save.name}_last.safetensors"), dtype=save_weight_dtype)
    log.finish()
    print("Done.")
    return losses


def dist_env(env=None):
    """(rank, world size, local rank) as torchrun exports them; (0, 1, 0) for a plain `python -m` launch."""
    import os
    env = os.environ if env is None else env
    return int(env.get("RANK", "0")), int(env.get("WORLD_SIZE", "1")), int(env.get("LOCAL_RANK", "0"))


def main(args=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--config_file", required=True, help="Config file for training.")
    ap.add_argument("--xl", action="store_true", help="SDXL loop (train_lora_xl.py)")
    a = ap.parse_args(args)
    config = config_util.load_config_from_yaml(a.config_file)
    prompts = config_util.load_prompts_from_yaml(config.prompts_file)
    # data parallel: `python -m torch.distributed.run --nproc-per-node N -m leco_b200.train_lora ...` —
    # one process per GPU, each prompt's batch_size is the GLOBAL batch and is split over the ranks
    rank, world, local = dist_env()
    device = torch.device("cuda", local)
    torch.cuda.set_device(device)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=device)
    try:
        train(config, prompts, xl=a.xl, device=device, rank=rank, world_size=world)
    finally:
        if world > 1:
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
