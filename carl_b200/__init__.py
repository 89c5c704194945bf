"""carl_b200 -- B200-native batched-step engine behind CARL's contextual-env API.

Import surface mirrors the reference for the hot path: ``carl_b200.context`` ↔ ``carl.context``,
``carl_b200.envs`` ↔ ``carl.envs``. Importing the package does not need a GPU; constructing an
env does (libcarlb has no CPU fallback).
"""
__version__ = "0.1.0"
