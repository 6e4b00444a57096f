"""Name-only stand-in for ``diffusers`` (NOT installed in this image).

The reference's lora.py:11, train_util.py:6 and model_util.py:5-16 import these
names purely as type annotations / constructors that the oracle replaces; this
stub lets the reference's own unmodified files import so the oracle can be
pinned against them (oracle/ref_loader.py).  Test infrastructure only.
"""
from oracle.unet_ref import UNet2DConditionModel  # noqa: F401


class SchedulerMixin:  # annotation only
    pass


class StableDiffusionPipeline:  # never constructed by the oracle
    pass


class StableDiffusionXLPipeline:
    pass
