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
    CARLBraxPusher,
    CARLBraxReacher,
    CARLBraxWalker2d,
)
from carl_b200.envs.mixed import MixedBatch  # noqa: F401,E402
