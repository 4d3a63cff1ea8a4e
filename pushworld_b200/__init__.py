"""pushworld_b200 -- B200-native batched PushWorld dynamics engine.

Drop-in for ONE hot path of google-deepmind/pushworld: ``PushWorldPuzzle.get_next_state`` +
goal check / reward (+ RGB observation), behind the reference's own ``step()/reset()`` surface.
See DESIGN.md for the architecture and INTEGRATION.md for the C-ABI boundary.
"""

__version__ = "0.1.0"

from . import tables  # noqa: F401  (pure host logic, importable without CUDA)

__all__ = ["tables", "Engine", "PushWorldPuzzle", "PushWorldVectorEnv"]


def __getattr__(name):
    # CUDA-facing modules import torch and the native library lazily
    if name == "Engine":
        from .engine import Engine

        return Engine
    if name == "PushWorldPuzzle":
        from .puzzle import PushWorldPuzzle

        return PushWorldPuzzle
    if name == "PushWorldVectorEnv":
        from .vector_env import PushWorldVectorEnv

        return PushWorldVectorEnv
    raise AttributeError(name)
