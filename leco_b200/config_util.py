"""Config / prompt settings with the reference's field names and defaults (config_util.py:14-104, prompt_util.py:43-67).

The reference's own `config_util.RootConfig` / `prompt_util.PromptSettings` objects (pydantic) can be handed to
`leco_b200.train_lora.train` unchanged — it only reads attributes.  These plain dataclasses exist so that a YAML written
for the reference (examples/config.yaml, examples/prompts.yaml) also loads where the reference's files are absent (the
GPU box): same keys, same defaults, same ValueErrors."""
from __future__ import annotations

from dataclasses import dataclass, field, fields
from typing import List, Optional

import torch
import yaml

PRECISION_TYPES = ("fp32", "fp16", "bf16", "float32", "float16", "bfloat16")
NETWORK_TYPES = ("lierla", "c3lier")
SCHEDULERS = ("ddim", "ddpm", "lms", "euler_a")
ACTION_TYPES = ("erase", "enhance")


def _build(cls, values: Optional[dict]):
    values = dict(values or {})
    known = {f.name for f in fields(cls)}
    # pydantic's BaseModel ignores unknown keys (examples/xl_config.yaml carries a stray train.batch_size): so do we
    return cls(**{k: v for k, v in values.items() if k in known})


@dataclass
class PretrainedModelConfig:
    name_or_path: str
    v2: bool = False
    v_pred: bool = False
    clip_skip: Optional[int] = None


@dataclass
class NetworkConfig:
    type: str = "lierla"
    rank: int = 4
    alpha: float = 1.0
    training_method: str = "full"

    def __post_init__(self):
        from .lora import TRAINING_METHODS
        if self.type not in NETWORK_TYPES:
            raise ValueError(f"network.type must be one of {NETWORK_TYPES}")
        if self.training_method not in TRAINING_METHODS:
            raise ValueError(f"network.training_method must be one of {TRAINING_METHODS}")
        self.rank, self.alpha = int(self.rank), float(self.alpha)


@dataclass
class TrainConfig:
    precision: str = "bfloat16"
    noise_scheduler: str = "ddim"
    iterations: int = 500
    lr: float = 1e-4
    optimizer: str = "adamw"
    optimizer_args: str = ""
    lr_scheduler: str = "constant"
    max_denoising_steps: int = 50

    def __post_init__(self):
        if self.precision not in PRECISION_TYPES:
            raise ValueError(f"Invalid precision type: {self.precision}")
        if self.noise_scheduler not in SCHEDULERS:
            raise ValueError(f"train.noise_scheduler must be one of {SCHEDULERS}")
        self.lr = float(self.lr)            # `lr: 1e-4` is a YAML string; pydantic coerces it, so do we
        self.iterations, self.max_denoising_steps = int(self.iterations), int(self.max_denoising_steps)


@dataclass
class SaveConfig:
    name: str = "untitled"
    path: str = "./output"
    per_steps: int = 200
    precision: str = "float32"


@dataclass
class LoggingConfig:
    use_wandb: bool = False
    verbose: bool = False


@dataclass
class OtherConfig:
    use_xformers: bool = False


@dataclass
class RootConfig:
    prompts_file: str
    pretrained_model: PretrainedModelConfig
    network: NetworkConfig
    train: TrainConfig = field(default_factory=TrainConfig)
    save: SaveConfig = field(default_factory=SaveConfig)
    logging: LoggingConfig = field(default_factory=LoggingConfig)
    other: OtherConfig = field(default_factory=OtherConfig)


def parse_precision(precision: str) -> torch.dtype:
    """config_util.py:75-83."""
    if precision in ("fp32", "float32"):
        return torch.float32
    if precision in ("fp16", "float16"):
        return torch.float16
    if precision in ("bf16", "bfloat16"):
        return torch.bfloat16
    raise ValueError(f"Invalid precision type: {precision}")


def load_config_from_yaml(config_path: str) -> RootConfig:
    """config_util.py:86-104."""
    with open(config_path, "r") as f:
        cfg = yaml.load(f, Loader=yaml.FullLoader)
    if "prompts_file" not in cfg or "pretrained_model" not in cfg or "network" not in cfg:
        raise ValueError("config needs prompts_file, pretrained_model and network sections")
    return RootConfig(prompts_file=cfg["prompts_file"],
                      pretrained_model=_build(PretrainedModelConfig, cfg["pretrained_model"]),
                      network=_build(NetworkConfig, cfg["network"]), train=_build(TrainConfig, cfg.get("train")),
                      save=_build(SaveConfig, cfg.get("save")), logging=_build(LoggingConfig, cfg.get("logging")),
                      other=_build(OtherConfig, cfg.get("other")))


@dataclass
class PromptSettings:
    """prompt_util.py:43-67 incl. the defaulting validator (positive <- target, neutral <- unconditional)."""
    target: str
    positive: Optional[str] = None
    unconditional: str = ""
    neutral: Optional[str] = None
    action: str = "erase"
    guidance_scale: float = 1.0
    resolution: int = 512
    dynamic_resolution: bool = False
    batch_size: int = 1
    dynamic_crops: bool = False

    def __post_init__(self):
        if self.positive is None:
            self.positive = self.target
        if self.neutral is None:
            self.neutral = self.unconditional
        if self.action not in ACTION_TYPES:
            raise ValueError("action must be erase or enhance")


def load_prompts_from_yaml(path: str) -> List[PromptSettings]:
    """prompt_util.py:151-160."""
    with open(path, "r") as f:
        prompts = yaml.safe_load(f)
    if not prompts:
        raise ValueError("prompts file is empty")
    out = []
    for p in prompts:
        if "target" not in p:
            raise ValueError("target must be specified")
        out.append(_build(PromptSettings, p))
    return out
