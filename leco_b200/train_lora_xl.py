"""`python -m leco_b200.train_lora_xl --config_file <yaml>` — the reference's SDXL driver (train_lora_xl.py:37-385) on
the fused trainer.  The XL loop differs from train_lora.py only in what a prompt is (text + pooled embedding of two
encoders, train_util.py:106-130) and in the `add_time_ids` conditioning drawn every iteration (train_lora_xl.py:183-201);
both live in `LecoTrainer`, so this is `leco_b200.train_lora.train(..., xl=True)`."""
from __future__ import annotations

from . import train_lora
from .train_lora import DEVICE_CUDA, NUM_IMAGES_PER_PROMPT, flush  # noqa: F401  (the reference script's module names)


def train(config, prompts, **kw):
    return train_lora.train(config, prompts, xl=True, **kw)


def main(args=None):
    import sys
    argv = list(sys.argv[1:] if args is None else args)
    if "--xl" not in argv:
        argv.append("--xl")
    train_lora.main(argv)


if __name__ == "__main__":
    main()
