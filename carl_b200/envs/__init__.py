"""Env classes of the hot path (reference: ``carl/envs/__init__.py:7-113`` import names)."""
from carl_b200.envs.carl_env import CARLEnv, ContextTable  # noqa: F401
from carl_b200.envs.classic_control import (  # noqa: F401
    CARLAcrobot,
    CARLCartPole,
    CARLGymnasiumEnv,
    CARLMountainCar,
    CARLMountainCarContinuous,
    CARLPendulum,
)
from carl_b200.envs.brax import (  # noqa: F401,E402
    CARLBraxAnt,
    CARLBraxEnv,
    CARLBraxHalfcheetah,
    CARLBraxHopper,
    CARLBraxHumanoid,
    CARLBraxHumanoidStandup,
    CARLBraxInvertedDoublePendulum,
    CARLBraxInvertedPendulum,
    CARLBraxReacher,
    CARLBraxWalker2d,
)
from carl_b200.envs.mixed import MixedBatch  # noqa: F401,E402


def __getattr__(name: str):
    """The Brax body of ``carl/envs/brax/__init__.py`` that this engine does not build fails loudly
    (nothing is substituted): see DESIGN.md (f)."""
    from carl_b200.envs.brax import UNSUPPORTED_BODIES

    if name in UNSUPPORTED_BODIES:
        raise NotImplementedError(
            f"{name} is not built by carl_b200: the pusher needs body-vs-body contacts (DESIGN.md (f)). "
            "No fallback is substituted.")
    raise AttributeError(f"module 'carl_b200.envs' has no attribute {name!r}")
