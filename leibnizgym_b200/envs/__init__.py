"""Same import path shape as the reference (`leibnizgym.envs`)."""
from ..env import IsaacEnvBase, TrifingerEnv

__all__ = ["IsaacEnvBase", "TrifingerEnv"]
