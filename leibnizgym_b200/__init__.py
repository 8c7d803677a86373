"""leibnizgym_b200 — the per-step TriFinger MDP hot path of pairlab/leibnizgym
(reward + observation/state + reset) as hand-written sm_100a CUDA kernels behind
the reference's TrifingerEnv / VecTask interface.  See DESIGN.md."""
from .config import default_sim_config, default_trifinger_config, difficulty_config, resolve_config

__version__ = "0.1.0"
__all__ = ["TrifingerEnv", "IsaacEnvBase", "VecTask", "VecTaskPython", "difficulty_config", "resolve_config",
           "default_sim_config", "default_trifinger_config"]


def __getattr__(name):  # torch-importing modules load lazily so `import leibnizgym_b200` stays cheap
    if name in ("TrifingerEnv", "IsaacEnvBase"):
        from . import env
        return getattr(env, name)
    if name in ("VecTask", "VecTaskPython"):
        from .wrappers import vec_task
        return getattr(vec_task, name)
    raise AttributeError(name)
