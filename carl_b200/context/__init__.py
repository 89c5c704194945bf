"""Context plugin surface (reference: ``carl/context/``)."""
from carl_b200.context.context_space import (  # noqa: F401
    CategoricalContextFeature,
    ContextFeature,
    ContextSpace,
    NormalFloatContextFeature,
    NumericalContextFeature,
    UniformFloatContextFeature,
    UniformIntegerContextFeature,
)
from carl_b200.context.sampler import ContextSampler  # noqa: F401
from carl_b200.context.selection import (  # noqa: F401
    AbstractSelector,
    CustomSelector,
    RandomSelector,
    RoundRobinSelector,
    StaticSelector,
)
