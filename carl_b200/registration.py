"""gymnasium registration of the batched envs (reference: ``carl/__init__.py:26-87`` registers
``carl/<Env>-v0`` ids). gymnasium is optional: without it this is a no-op that returns []."""
from __future__ import annotations

ENV_NAMES = ["CARLCartPole", "CARLPendulum", "CARLAcrobot", "CARLMountainCar", "CARLMountainCarContinuous",
             "CARLBraxAnt", "CARLBraxHalfcheetah", "CARLBraxHopper", "CARLBraxWalker2d",
             "CARLBraxHumanoid", "CARLBraxHumanoidStandup", "CARLBraxInvertedPendulum", "CARLBraxInvertedDoublePendulum",
             "CARLBraxReacher", "CARLBraxPusher"]


def register_envs(namespace: str = "carl_b200") -> list[str]:
    try:
        import gymnasium
    except Exception:
        return []
    ids = []
    for name in ENV_NAMES:
        env_id = f"{namespace}/{name}-v0"
        if env_id not in gymnasium.registry:
            gymnasium.register(id=env_id, entry_point=f"carl_b200.envs:{name}", disable_env_checker=True,
                               order_enforce=False)
        ids.append(env_id)
    return ids
